"""qr::cacqr -- host-side mirror of the reference's CholeskyQR2 entry points (src/alg/qr/cacqr/cacqr.h:42-49).

    ci = cholinv.info(complete_inv, split, bc_mult_dim, 'U')
    args = cacqr.info(num_iter, ci)                 # 1 = CholeskyQR, 2 = CholeskyQR2   (cacqr.h:28-32)
    cacqr.factor(A, args, topo)                     # cacqr.hpp:217-248 -> args.Q (rect), args.R (packed upper)
"""
from __future__ import annotations
import ctypes as C
import torch
from . import _lib
from . import cholinv as _ci
from .matrix import matrix


class info:
    def __init__(self, num_iter: int, cholesky_inverse_args: _ci.info, serialize: bool = True):
        self.num_iter = int(num_iter)
        self.cholesky_inverse_args = cholesky_inverse_args
        self.serialize = serialize
        self.Q = None
        self.R = None
        self.n = 0


def factor(A: matrix, args: info, topo):
    ctx = topo.context()
    m, n = A.num_rows_global, A.num_columns_global
    dev = A.data.device
    pin = dev.type == "cpu"
    if args.Q is None or args.Q.numel() != A.data.numel() or args.Q.device != dev:
        args.Q = torch.empty_like(A.data, pin_memory=pin) if pin else torch.empty_like(A.data)
    lc = A.num_columns_local
    rcount = lc * (lc + 1) // 2 if args.serialize else lc * lc
    if args.R is None or args.R.numel() != rcount or args.R.device != dev:
        args.R = torch.empty(rcount, dtype=torch.float64, device=dev, pin_memory=pin)
    args.n = lc
    args.rows_local = A.num_rows_local
    cargs = args.cholesky_inverse_args._c()
    ctx.check(_lib.lib().capital_cacqr_factor_f64(ctx.handle, A.data.data_ptr(), m, n, args.num_iter, C.byref(cargs),
                                                  _lib.UPPERTRI_PACKED if args.serialize else _lib.RECT,
                                                  args.Q.data_ptr(), args.R.data_ptr()))


def construct_Q(args: info, topo=None) -> torch.Tensor:
    return args.Q.view(args.n, args.rows_local).t()


def construct_R(args: info, topo=None) -> torch.Tensor:
    return _ci._expand(args.R, args.n, args.serialize)


def validate(A: matrix, args: info, topo):
    """(residual, orthogonality) of qr::validate (test/qr/validate.hpp:7-52)."""
    ctx = topo.context()
    r, o = C.c_double(), C.c_double()
    ctx.check(_lib.lib().capital_cacqr_residual_f64(ctx.handle, A.data.data_ptr(), A.num_rows_global, A.num_columns_global,
                                                    args.Q.data_ptr(), _lib.UPPERTRI_PACKED if args.serialize else _lib.RECT,
                                                    args.R.data_ptr(), C.byref(r), C.byref(o)))
    return float(r.value), float(o.value)
