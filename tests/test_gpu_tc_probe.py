"""Pin the tcgen05 descriptor / TMEM layouts (tests/tc_layouts.py == csrc/tc_common.cuh) against numpy on the GPU.
Inputs are small integers, so every product and sum is exact in tf32 x tf32 -> fp32: results must be EQUAL."""
import ctypes as C

import numpy as np
import pytest
import torch

import tc_layouts as TL

pytestmark = pytest.mark.gpu


def run_probe(image_words, mmas, read_cols):
    from rl_replicas_b200 import _lib
    lib = _lib.load()
    img = torch.from_numpy(image_words.astype(np.uint32).view(np.int32)).cuda()
    rec = np.zeros(len(mmas), dtype=[("a", "<u8"), ("b", "<u8"), ("i", "<u4"), ("d", "<u4"), ("acc", "<u4"), ("kind", "<u4")])
    for k, m in enumerate(mmas):
        rec[k] = m if len(m) == 6 else m + (0,)
    mm = torch.from_numpy(rec.view(np.uint8)).cuda()
    out = torch.zeros(128 * read_cols, dtype=torch.float32, device="cuda")
    _lib.check(lib.b200rl_tc_probe(C.c_void_p(img.data_ptr()), img.numel(), C.c_void_p(mm.data_ptr()), len(mmas),
                                   read_cols, C.c_void_p(out.data_ptr()), int(torch.cuda.current_stream().cuda_stream)),
               "tc_probe")
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(128, read_cols)


def ints(rng, shape):
    return rng.integers(-4, 5, size=shape).astype(np.float32)


def layout_image(parts):
    """parts: list of uint32 arrays -> (image, [byte offsets]) with every part 1024-byte aligned."""
    offs, words = [], []
    pos = 0
    for p in parts:
        offs.append(pos)
        words.append(p)
        pos += p.size * 4
        pad = (-pos) % 1024
        if pad:
            words.append(np.zeros(pad // 4, dtype=np.uint32))
            pos += pad
    return np.concatenate(words), offs


def expect(D, M, N, cols):
    out = np.zeros((128, cols), dtype=np.float32)
    for m in range(M):
        out[TL.tmem_lane(m, M), :N] = D[m]
    return out


# kind::tf32: only K-major operands are pinned.  MN-major tf32 needs the special SWIZZLE_128B_BASE32B layout
# (plain SWIZZLE_128B + b_major = 1 yields zeros on B200 -- measured, tools/tc_decode.py), so one activation buffer
# cannot be read both ways in tf32; the product kernel uses bf16 splits (below), where it can.
@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [
    (128, 64, 64, 0, 0), (128, 16, 64, 0, 0), (128, 64, 32, 0, 0), (128, 8, 64, 0, 0),
])
def test_single_gemm(M, N, K, a_mn, b_mn):
    rng = np.random.default_rng(M * 1000 + N * 10 + K + a_mn * 3 + b_mn)
    A, B = ints(rng, (M, K)), ints(rng, (N, K))
    a_buf = TL.pack(A.T.copy()) if a_mn else TL.pack(A)          # MN-major: stored [K rows][M cols]
    b_store = B.T.copy() if b_mn else B
    if b_store.shape[0] % 8:
        b_store = np.concatenate([b_store, np.zeros((8 - b_store.shape[0] % 8, b_store.shape[1]), np.float32)])
    b_buf = TL.pack(b_store)
    image, (oa, ob) = layout_image([a_buf, b_buf])
    idesc = TL.idesc_tf32(M, N, a_mn, b_mn)
    mmas = []
    for j in range(K // 8):
        ad = TL.desc_mnmajor(oa, K, j) if a_mn else TL.desc_kmajor(oa, M, j)
        bd = TL.desc_mnmajor(ob, K, j) if b_mn else TL.desc_kmajor(ob, b_store.shape[0], j)
        mmas.append((ad, bd, idesc, 0, 1 if j else 0))
    cols = max(8, ((N + 7) // 8) * 8)
    got = run_probe(image, mmas, cols)
    np.testing.assert_array_equal(got, expect(A @ B.T, M, N, cols))


def test_accumulate_chain_and_column_offset():
    """two GEMMs into different TMEM column ranges + accumulation across separately issued chains"""
    rng = np.random.default_rng(1)
    A1, B1, A2, B2 = ints(rng, (128, 32)), ints(rng, (64, 32)), ints(rng, (128, 32)), ints(rng, (16, 32))
    image, (o1, o2, o3, o4) = layout_image([TL.pack(A1), TL.pack(B1), TL.pack(A2), TL.pack(B2)])
    i64, i16 = TL.idesc_tf32(128, 64, 0, 0), TL.idesc_tf32(128, 16, 0, 0)
    mmas = []
    for rep in range(2):  # D1 = 2 * A1 B1^T through accumulation
        for j in range(4):
            mmas.append((TL.desc_kmajor(o1, 128, j), TL.desc_kmajor(o2, 64, j), i64, 0, 1 if (j or rep) else 0))
    for j in range(4):
        mmas.append((TL.desc_kmajor(o3, 128, j), TL.desc_kmajor(o4, 16, j), i16, 64, 1 if j else 0))
    got = run_probe(image, mmas, 80)
    want = np.zeros((128, 80), np.float32)
    want[:, :64] = 2 * (A1 @ B1.T)
    want[:, 64:80] = A2 @ B2.T
    np.testing.assert_array_equal(got, want)


def test_3xtf32_split_reaches_fp32_accuracy():
    """error-compensated split a = hi + lo (both tf32-representable): hi*hi + hi*lo + lo*hi ~ fp32 product"""
    rng = np.random.default_rng(2)
    A, B = rng.standard_normal((128, 64)).astype(np.float32), rng.standard_normal((64, 64)).astype(np.float32)

    def split(x):
        hi = (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        return hi, (x - hi).astype(np.float32)

    ah, al = split(A)
    bh, bl = split(B)
    image, (oah, oal, obh, obl) = layout_image([TL.pack(ah), TL.pack(al), TL.pack(bh), TL.pack(bl)])
    idesc = TL.idesc_tf32(128, 64, 0, 0)
    mmas = []
    for (oa, ob) in ((oal, obh), (oah, obl), (oah, obh)):  # small terms first
        for j in range(8):
            mmas.append((TL.desc_kmajor(oa, 128, j), TL.desc_kmajor(ob, 64, j), idesc, 0, 1 if mmas else 0))
    got = run_probe(image, mmas, 64)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    err = np.abs(got - ref).max() / np.abs(ref).max()
    one = np.abs((ah.astype(np.float64) @ bh.astype(np.float64).T) - ref).max() / np.abs(ref).max()
    print("3xTF32 rel err", err, " 1xTF32 rel err", one)
    assert err < 2e-6 and one > 1e-4


# ----------------------------------------------- kind::f16, bf16 operands -----------------------------------------
@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [
    (128, 64, 64, 0, 0), (128, 64, 64, 0, 1), (128, 64, 64, 1, 0), (128, 16, 64, 0, 0), (128, 64, 16, 0, 1),
    (64, 64, 128, 1, 1), (64, 8, 128, 1, 1), (64, 32, 128, 1, 1), (64, 24, 128, 1, 1), (128, 64, 32, 0, 0),
    (128, 32, 64, 0, 1), (128, 64, 64, 1, 1), (64, 16, 128, 1, 1), (64, 64, 64, 1, 1),
])
def test_single_gemm_bf16(M, N, K, a_mn, b_mn):
    rng = np.random.default_rng(M * 1000 + N * 10 + K + a_mn * 3 + b_mn + 7)
    A, B = ints(rng, (M, K)), ints(rng, (N, K))
    bits = TL.f32_to_bf16_bits
    a_buf = TL.pack16(bits(A.T.copy())) if a_mn else TL.pack16(bits(A))
    b_store = B.T.copy() if b_mn else B
    if b_store.shape[0] % 8:
        b_store = np.concatenate([b_store, np.zeros((8 - b_store.shape[0] % 8, b_store.shape[1]), np.float32)])
    b_buf = TL.pack16(bits(b_store))
    image, (oa, ob) = layout_image([a_buf, b_buf])
    idesc = TL.idesc_bf16(M, N, a_mn, b_mn)
    mmas = []
    for j in range(K // 16):
        ad = TL.desc16_mnmajor(oa, K, j) if a_mn else TL.desc16_kmajor(oa, M, j)
        bd = TL.desc16_mnmajor(ob, K, j) if b_mn else TL.desc16_kmajor(ob, b_store.shape[0], j)
        mmas.append((ad, bd, idesc, 0, 1 if j else 0, 1))
    cols = max(8, ((N + 7) // 8) * 8)
    got = run_probe(image, mmas, cols)
    np.testing.assert_array_equal(got, expect(A @ B.T, M, N, cols))


def test_bf16_three_way_split_reaches_fp32_accuracy():
    """x = h + m + l (three bf16, 24 mantissa bits): hh + hm + mh + hl + lh + mm ~ fp32 product (6 MMAs)"""
    rng = np.random.default_rng(3)
    A, B = rng.standard_normal((128, 64)).astype(np.float32), rng.standard_normal((64, 64)).astype(np.float32)

    def split3(x):
        parts, r = [], x.astype(np.float32)
        for _ in range(3):
            b = TL.f32_to_bf16_bits(r)
            parts.append(b)
            r = (r - TL.bf16_bits_to_f32(b)).astype(np.float32)
        return parts

    a3, b3 = split3(A), split3(B)
    bufs = [TL.pack16(x) for x in a3 + b3]
    image, offs = layout_image(bufs)
    oa, ob = offs[:3], offs[3:]
    idesc = TL.idesc_bf16(128, 64, 0, 0)
    mmas = []
    for (i, j) in ((1, 1), (0, 2), (2, 0), (0, 1), (1, 0), (0, 0)):  # small terms first
        for k in range(4):
            mmas.append((TL.desc16_kmajor(oa[i], 128, k), TL.desc16_kmajor(ob[j], 64, k), idesc, 0, 1 if mmas else 0, 1))
    got = run_probe(image, mmas, 64)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    err = np.abs(got - ref).max() / np.abs(ref).max()
    three = [TL.bf16_bits_to_f32(a3[i]).astype(np.float64) @ TL.bf16_bits_to_f32(b3[j]).astype(np.float64).T
             for (i, j) in ((0, 0), (0, 1), (1, 0))]
    err3 = np.abs(sum(three) - ref).max() / np.abs(ref).max()
    print("bf16 6-term rel err", err, " 3-term rel err", err3)
    assert err < 1e-6


@pytest.mark.parametrize("M,N,col0,a_kmajor_k0", [(64, 16, 32, None), (64, 32, 0, None), (128, 64, 32, 2)])
def test_bf16_operand_at_column_offset(M, N, col0, a_kmajor_k0):
    """One 64-column block holds X (cols 0..31) and dOut (cols 32..47): read a sub-range of columns
    (a) MN-major as the B operand of dW^T (start address + col0*2 bytes), (b) K-major as the A operand (k-step 2)."""
    rng = np.random.default_rng(M + N + col0)
    bits = TL.f32_to_bf16_bits
    R = 128
    XD = ints(rng, (R, 64))
    if a_kmajor_k0 is None:
        Aact = ints(rng, (R, M))                       # e.g. H2 [rows][features], MN-major A (M = features)
        image, (oa, ob) = layout_image([TL.pack16(bits(Aact)), TL.pack16(bits(XD))])
        idesc = TL.idesc_bf16(M, N, 1, 1)
        mmas = [(TL.desc16_mnmajor(oa, R, j), TL.desc16_mnmajor(ob + col0 * 2, R, j), idesc, 0, 1 if j else 0, 1)
                for j in range(R // 16)]
        want = Aact.T @ XD[:, col0:col0 + N]
    else:
        W = ints(rng, (16, 64))                        # W3 [o][i]: B MN-major, K = o (16 rows), N = i
        image, (oa, ob) = layout_image([TL.pack16(bits(XD)), TL.pack16(bits(W))])
        idesc = TL.idesc_bf16(128, 64, 0, 1)
        mmas = [(TL.desc16_kmajor(oa, R, a_kmajor_k0), TL.desc16_mnmajor(ob, 16, 0), idesc, 0, 0, 1)]
        want = XD[:, col0:col0 + 16] @ W
    cols = max(8, N)
    got = run_probe(image, mmas, cols)
    np.testing.assert_array_equal(got, expect(want, M, N, cols))
