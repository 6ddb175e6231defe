"""Where the tolerances of tests/test_gpu_zz_late.py come from: a numpy emulation of the mixed-precision trailing update
(operands rounded to TF32 with round-to-nearest-away as `cvt.rna.tf32.f32` does, FP32 accumulation, optional hi + lo split) inside a
plain recursive Cholesky on the reference's generator.  CPU only; python tools/tf32_emulate.py [n ...]"""
import os
import sys

import numpy as np
import scipy.linalg as sl

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import capital_oracle as co  # noqa: E402


def tf32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000  # drop 13 mantissa bits, ties away from zero
    return u.astype(np.uint32).view(np.float32)


def product(A, B, passes):
    """A^T B for A: k x m, B: k x n (FP64 in, FP64 out) the way gemm_tf32.cu computes it"""
    ah, bh = tf32(A), tf32(B)
    if passes == 1:
        return (ah.T @ bh).astype(np.float64)
    al, bl = tf32(A - ah.astype(np.float64)), tf32(B - bh.astype(np.float64))
    return (ah.T @ bh + ah.T @ bl + al.T @ bh).astype(np.float64)


def cholesky(a, bc, passes, min_k):
    n = a.shape[0]
    if n <= bc:
        return np.linalg.cholesky(a).T
    s = n // 2
    R11 = cholesky(a[:s, :s], bc, passes, min_k)
    R12 = sl.solve_triangular(R11, a[:s, s:], trans="T", lower=False)
    S = a[s:, s:] - (product(R12, R12, passes) if passes and s >= min_k else R12.T @ R12)
    R = np.zeros_like(a)
    R[:s, :s], R[:s, s:], R[s:, s:] = R11, R12, cholesky(S, bc, passes, min_k)
    return R


if __name__ == "__main__":
    for n in [int(v) for v in sys.argv[1:]] or [1024, 4096]:
        a = co.spd_global(n)
        print(f"n={n}: " + ", ".join(f"{name} residual {co.cholesky_residual(a, cholesky(a, 64, p, 256)):.2e}"
                                     for name, p in (("f64", 0), ("tf32", 1), ("tf32x3", 3))))
    rng = np.random.default_rng(0)
    for k in (512, 4096):
        A, B = rng.standard_normal((k, 256)), rng.standard_normal((k, 384))
        ref, den = A.T @ B, (np.abs(A).T @ np.abs(B)).max()
        print(f"product k={k}: " + ", ".join(f"{p} pass(es) {np.abs(product(A, B, p) - ref).max() / den:.2e}" for p in (1, 3)))
