"""matmult::summa -- host-side mirror of the reference's SUMMA GEMM entry point (src/alg/matmult/summa/summa.h:24-26) in the
T*N form:  summa.invoke(A, B, C, topo, alpha, beta)  computes  C = alpha * A^T B + beta * C  on the process grid, with
A (k x m), B (k x n), C (m x n) element-cyclic `matrix` objects (what the validators call: test/cholesky/validate.hpp:35)."""
from __future__ import annotations
from . import _lib
from .matrix import matrix


def invoke(A: matrix, B: matrix, C: matrix, topo, alpha: float = 1.0, beta: float = 0.0):
    k, m = A.num_rows_global, A.num_columns_global
    n = B.num_columns_global
    if B.num_rows_global != k or C.num_rows_global != m or C.num_columns_global != n:
        raise ValueError("summa.invoke: need A (k x m), B (k x n), C (m x n)")
    ctx = topo.context()
    ctx.check(_lib.lib().capital_summa_gemm_tn_f64(ctx.handle, m, n, k, alpha, A.data.data_ptr(), B.data.data_ptr(), beta, C.data.data_ptr()))
