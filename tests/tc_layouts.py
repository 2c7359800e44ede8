"""numpy model of the shared-memory / TMEM layouts used by the tcgen05 kernels (mirrors csrc/tc_common.cuh).

One buffer format serves every operand: a [R rows][C cols] fp32 array stored as column blocks of 32 floats
(128 bytes per row inside a block, rows at a 128-byte pitch, blocks R*128 bytes apart) with the SWIZZLE_128B XOR
(16-byte chunk index ^= row % 8).  Read with a K-major descriptor it is an operand whose K runs along the columns;
read with an MN-major descriptor it is an operand whose K runs along the rows.
"""
import numpy as np


def buf_bytes(R, C):
    return ((C + 31) // 32) * R * 128


def buf_offset(r, c, R):
    block, cc = c // 32, c % 32
    chunk = (cc * 4) // 16
    return block * (R * 128) + r * 128 + ((chunk ^ (r % 8)) * 16) + (cc * 4) % 16


def pack(mat):
    """[R,C] float32 -> uint32 words of the swizzled buffer (R multiple of 8)."""
    R, C = mat.shape
    assert R % 8 == 0
    out = np.zeros(buf_bytes(R, C) // 4, dtype=np.uint32)
    bits = np.ascontiguousarray(mat, dtype=np.float32).view(np.uint32)
    for r in range(R):
        for c in range(C):
            out[buf_offset(r, c, R) // 4] = bits[r, c]
    return out


def smem_desc(addr, lbo, sbo):
    return ((addr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (2 << 61)


def idesc_tf32(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def desc_kmajor(base, R, kstep):
    """operand [R rows = M|N index][cols = K]; one MMA consumes 8 K-elements = 32 bytes of a row."""
    return smem_desc(base + (kstep // 4) * (R * 128) + (kstep % 4) * 32, 16, 1024)


def desc_mnmajor(base, R, kstep, lbo=None, sbo=1024):
    """operand [R rows = K index][cols = M|N index]; one MMA consumes 8 rows = 1024 bytes; column blocks (32 MN
    elements) are R*128 bytes apart (leading byte offset)."""
    return smem_desc(base + kstep * 1024, R * 128 if lbo is None else lbo, sbo)


def tmem_lane(m, M):
    return m if M == 128 else (m % 16) + 32 * (m // 16)


# ---- 16-bit (bf16) operands: column blocks of 64 elements (128 bytes per row), same SWIZZLE_128B XOR -------------
def f32_to_bf16_bits(x):
    """round-to-nearest-even float32 -> bf16 bit patterns (uint16)"""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def bf16_bits_to_f32(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def buf16_bytes(R, C):
    return ((C + 63) // 64) * R * 128


def buf16_offset(r, c, R):
    block, cc = c // 64, c % 64
    chunk = (cc * 2) // 16
    return block * (R * 128) + r * 128 + ((chunk ^ (r % 8)) * 16) + (cc * 2) % 16


def pack16(mat_bits):
    """[R,C] uint16 bf16 bit patterns -> uint32 words of the swizzled buffer"""
    R, C = mat_bits.shape
    assert R % 8 == 0
    out = np.zeros(buf16_bytes(R, C) // 2, dtype=np.uint16)
    for r in range(R):
        for c in range(C):
            out[buf16_offset(r, c, R) // 2] = mat_bits[r, c]
    return out.view(np.uint32)


def idesc_bf16(M, N, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def desc16_kmajor(base, R, kstep):
    """operand [R rows = M|N][cols = K] bf16; one MMA consumes 16 K-elements = 32 bytes of a row"""
    return smem_desc(base + (kstep // 4) * (R * 128) + (kstep % 4) * 32, 16, 1024)


def desc16_mnmajor(base, R, kstep, lbo=None, sbo=1024):
    """operand [R rows = K][cols = M|N] bf16; one MMA consumes 16 rows = 2048 bytes; 64-element column blocks are
    R*128 bytes apart"""
    return smem_desc(base + kstep * 2048, R * 128 if lbo is None else lbo, sbo)
