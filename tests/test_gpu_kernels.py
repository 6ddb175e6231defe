"""GPU parity of the leaf kernels through the C ABI: DMMA GEMM (every structure flag, ragged and unaligned
windows), fused potrf+trtri, and the generators (bit-exact against the oracle restatement)."""
import numpy as np
import pytest
import torch
import capital_b200 as cb
from capital_b200 import _lib
from oracle import capital_oracle as co

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return cb.topo.square(1, 0, 1).context()


def colmajor(rows, cols, ld=None, seed=0):
    ld = ld or rows
    g = torch.Generator(device="cuda").manual_seed(seed)
    buf = torch.rand(ld * cols, dtype=torch.float64, device="cuda", generator=g) - 0.5
    return buf, buf.view(cols, ld).t()[:rows]  # (flat storage, rows x cols view)


FLAGS = [0, _lib.GEMM_A_UPPER, _lib.GEMM_B_UPPER, _lib.GEMM_B_LOWER, _lib.GEMM_C_UPPER, _lib.GEMM_A_LOWER | _lib.GEMM_B_UPPER | _lib.GEMM_C_UPPER,
         _lib.GEMM_A_UPPER | _lib.GEMM_B_UPPER | _lib.GEMM_C_UPPER]


@pytest.mark.parametrize("m,n,k", [(64, 64, 16), (128, 128, 64), (200, 136, 72), (1, 5, 3), (333, 1000, 517), (2048, 1536, 1024), (1280, 1280, 2000)])
@pytest.mark.parametrize("flags", FLAGS)
def test_gemm_tn_matches_fp64_reference(ctx, m, n, k, flags):
    tri = flags & (_lib.GEMM_A_UPPER | _lib.GEMM_A_LOWER | _lib.GEMM_B_UPPER | _lib.GEMM_B_LOWER)
    if tri and not (m == n == k or (flags & (_lib.GEMM_A_UPPER | _lib.GEMM_A_LOWER) and k == m and not flags & (_lib.GEMM_B_UPPER | _lib.GEMM_B_LOWER))
                    or (flags & (_lib.GEMM_B_UPPER | _lib.GEMM_B_LOWER) and k == n and not flags & (_lib.GEMM_A_UPPER | _lib.GEMM_A_LOWER))):
        k = m if flags & (_lib.GEMM_A_UPPER | _lib.GEMM_A_LOWER) else n
        if (flags & (_lib.GEMM_A_UPPER | _lib.GEMM_A_LOWER)) and (flags & (_lib.GEMM_B_UPPER | _lib.GEMM_B_LOWER)):
            n = m
    if (flags & _lib.GEMM_C_UPPER) and m != n:
        n = m
    lda, ldb, ldc = k + 6 + (k & 1), k + 2 + (k & 1), m + 3
    fa, A = colmajor(k, m, lda, 1)
    fb, B = colmajor(k, n, ldb, 2)
    fc, Cm = colmajor(m, n, ldc, 3)
    if flags & _lib.GEMM_A_UPPER: A.copy_(torch.triu(A))
    if flags & _lib.GEMM_A_LOWER: A.copy_(torch.tril(A))
    if flags & _lib.GEMM_B_UPPER: B.copy_(torch.triu(B))
    if flags & _lib.GEMM_B_LOWER: B.copy_(torch.tril(B))
    alpha, beta = -0.75, 0.5
    ref = alpha * (A.t() @ B) + beta * Cm
    if flags & _lib.GEMM_C_UPPER:
        ref = torch.where(torch.ones_like(ref, dtype=torch.bool).triu(), ref, Cm)
    ctx.check(_lib.lib().capital_blas_gemm_tn_f64(ctx.handle, m, n, k, alpha, fa.data_ptr(), lda, fb.data_ptr(), ldb, beta, fc.data_ptr(), ldc, flags))
    ctx.synchronize()
    err = (Cm - ref).abs().max().item()
    assert err <= 1e-13 * max(1.0, k ** 0.5) * 4, (m, n, k, flags, err)
    # padding rows of C (ld > m) must be untouched
    fresh, _ = colmajor(m, n, ldc, 3)
    assert torch.equal(fc.view(n, ldc)[:, m:], fresh.view(n, ldc)[:, m:])


def test_gemm_tn_unaligned_window(ctx):
    """operand windows starting at an odd row (8-byte but not 16-byte aligned): TMA map is built on the aligned base."""
    K, M, N, ld = 96, 70, 50, 200
    fa, A = colmajor(ld, M, ld, 4)
    fb, B = colmajor(ld, N, ld, 5)
    fc, Cm = colmajor(M, N, M, 6)
    for ra, rb in [(1, 0), (3, 7), (0, 5)]:
        Aw, Bw = A[ra:ra + K], B[rb:rb + K]
        ref = Aw.t() @ Bw
        ctx.check(_lib.lib().capital_blas_gemm_tn_f64(ctx.handle, M, N, K, 1.0, fa.data_ptr() + 8 * ra, ld, fb.data_ptr() + 8 * rb, ld, 0.0, fc.data_ptr(), M, 0))
        ctx.synchronize()
        assert (Cm - ref).abs().max().item() < 1e-12


@pytest.mark.parametrize("n", [1, 5, 33, 64, 65, 100, 128, 200, 512, 777, 1024])
def test_potrf_trtri_matches_lapack(ctx, n):
    a = torch.from_numpy(co.spd_global(n)).cuda()
    ld = n + (n & 1)
    R = torch.full((n, ld), 7.0, dtype=torch.float64, device="cuda")
    Ri = torch.full((n, ld), 7.0, dtype=torch.float64, device="cuda")
    acm = a.t().contiguous()  # symmetric anyway; column-major storage
    ctx.check(_lib.lib().capital_lapack_potrf_trtri_f64(ctx.handle, n, acm.data_ptr(), n, R.data_ptr(), ld, Ri.data_ptr(), ld))
    Rm, Rim = R.t()[:n], Ri.t()[:n]
    ref = torch.linalg.cholesky(a, upper=True)
    assert (Rm - ref).abs().max().item() < 1e-13 * n
    assert torch.equal(Rm.tril(-1), torch.zeros_like(Rm)) and torch.equal(Rim.tril(-1), torch.zeros_like(Rim))
    assert (Rim @ Rm - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max().item() < 1e-13


def test_not_spd_is_reported(ctx):
    n = 96
    a = torch.from_numpy(co.spd_global(n)).cuda()
    a[40, 40] = -5.0
    R = torch.empty(n * n, dtype=torch.float64, device="cuda")
    Ri = torch.empty_like(R)
    st = _lib.lib().capital_lapack_potrf_trtri_f64(ctx.handle, n, a.data_ptr(), n, R.data_ptr(), n, Ri.data_ptr(), n)
    assert st == _lib.ERR_NOT_SPD


@pytest.mark.parametrize("n,size,c", [(37, 1, 1), (64, 8, 2), (101, 8, 2), (300, 27, 3)])
def test_distribute_symmetric_bit_exact(n, size, c):
    for rank in (0, size - 1, size // 2):
        t = cb.topo.square(size, rank, c)
        g = t.grid
        one = _lib.Context(g, 0)  # generators need no communicator
        L = co.local_dim(n, t.d)
        out = torch.empty(L * L, dtype=torch.float64, device="cuda")
        one.check(_lib.lib().capital_distribute_symmetric_f64(one.handle, out.data_ptr(), n, 1))
        ref = co.spd_local(n, t.d, t.x, t.y)
        assert np.array_equal(out.cpu().numpy().reshape(L, L, order="F"), ref)
        one.close()


@pytest.mark.parametrize("m,n,size,c", [(64, 8, 1, 1), (1000, 32, 8, 1), (257, 33, 8, 2), (4096, 64, 4, 1)])
def test_distribute_random_bit_exact(m, n, size, c):
    for rank in (0, size - 1):
        t = cb.topo.rect(size, rank, c)
        one = _lib.Context(t.grid, 0)
        lr, lc = co.local_dim(m, t.d), co.local_dim(n, t.c)
        out = torch.empty(lr * lc, dtype=torch.float64, device="cuda")
        one.check(_lib.lib().capital_distribute_random_f64(one.handle, out.data_ptr(), m, n, rank // c))
        ref = co.random_local(m, n, t.c, t.d, t.x, t.y, rank // c)
        assert np.array_equal(out.cpu().numpy().reshape(lr, lc, order="F"), ref)
        one.close()
