"""Decode how tcgen05 addresses an MN-major SWIZZLE_128B operand: A = [I;0] K-major (known good), B raw buffer
filled with word indices, so D[k][n] = word index the hardware fetched for B(n,k)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import tc_layouts as TL
from test_gpu_tc_probe import run_probe, layout_image

def decode(N, K_rows, lbo, sbo, b_words, kstep=0, a_mn=0):
    A = np.zeros((128, 8), np.float32)
    for k in range(8):
        A[k, k] = 1.0
    a_buf = TL.pack(np.concatenate([A, np.zeros((128, 24), np.float32)], axis=1))
    res = []
    for part in (lambda w: w % 1024, lambda w: w // 1024):
        b_buf = np.array([part(w) for w in range(b_words)], dtype=np.float32).view(np.uint32)
        image, (oa, ob) = layout_image([a_buf, b_buf])
        idesc = TL.idesc_tf32(128, N, 0, 1)
        mmas = [(TL.desc_kmajor(oa, 128, 0), TL.smem_desc(ob + kstep * 1024, lbo, sbo), idesc, 0, 0)]
        res.append(run_probe(image, mmas, N))
    return (res[0] + 1024 * res[1]).astype(np.int64)  # [128 lanes][N]: row k (k<8) holds word index of B(n,k)

if __name__ == "__main__":
    N, R = 64, 64
    words = TL.buf_bytes(R, N) // 4
    for (lbo, sbo) in ((R * 128, 1024), (1024, R * 128), (16, 1024), (128, 1024)):
        d = decode(N, R, lbo, sbo, words)
        print(f"--- LBO={lbo} SBO={sbo}: word index fetched for B(n,k), k=0..7 rows, n=0..63")
        for k in range(8):
            print("k=%d:" % k, " ".join("%5d" % x for x in d[k, :64]))
        exp = np.array([[TL.buf_offset(k, n, R) // 4 for n in range(N)] for k in range(8)])
        print("matches my model:", np.array_equal(d[:8, :N], exp))
