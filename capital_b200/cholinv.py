"""cholesky::cholinv -- host-side mirror of the reference's entry points (src/alg/cholesky/cholinv/cholinv.h:46-53).

    args = cholinv.info(complete_inv, split, bc_mult_dim, 'U')      # cholinv.h:25-30
    cholinv.factor(A, args, topo)                                    # cholinv.hpp:6-28  -> args.R, args.Rinv
    R = cholinv.construct_R(args, topo)                              # cholinv.hpp:30-37 -> rect, zero lower

Outputs are packed upper-triangular local blocks (policy::cholinv::Serialize) unless serialize=False."""
from __future__ import annotations
import ctypes as C
import torch
from . import _lib
from .matrix import matrix


class info:
    def __init__(self, complete_inv, split: int, bc_mult_dim: int, dir: str = "U", serialize: bool = True):
        if split <= 0 or dir != "U":
            raise ValueError("cholinv requires split > 0 and dir == 'U' (cholinv.hpp:9)")
        self.complete_inv, self.split, self.bc_mult_dim, self.dir = int(bool(complete_inv)), int(split), int(bc_mult_dim), dir
        self.serialize = serialize
        self.R = None      # torch tensors: packed upper L(L+1)/2 (or L*L rect)
        self.Rinv = None
        self.local_dim = 0
        self.global_dim = 0

    def _c(self) -> _lib.CholinvArgs:
        return _lib.CholinvArgs(self.complete_inv, self.split, self.bc_mult_dim, self.dir.encode())


def _register(args: info, L: int, device):
    count = L * (L + 1) // 2 if args.serialize else L * L
    for name in ("R", "Rinv"):  # matrix::_register_: allocate on first use only (matrix.hpp:141-155)
        t = getattr(args, name)
        if t is None or t.numel() != count or t.device != device:
            pin = device.type == "cpu"
            setattr(args, name, torch.empty(count, dtype=torch.float64, device=device, pin_memory=pin))


def factor(A: matrix, args: info, topo):
    """A is never modified (const&).  Results land in args.R / args.Rinv, on the same device kind as A."""
    ctx = topo.context()
    n = A.num_rows_global
    assert A.num_columns_global == n
    L = A.num_rows_local
    _register(args, L, A.data.device)
    args.local_dim, args.global_dim = L, n
    cargs = args._c()
    st = _lib.lib().capital_cholinv_factor_f64(ctx.handle, A.data.data_ptr(), n, C.byref(cargs),
                                               _lib.UPPERTRI_PACKED if args.serialize else _lib.RECT,
                                               args.R.data_ptr(), args.Rinv.data_ptr())
    ctx.check(st)


def _expand(packed: torch.Tensor, L: int, serialize: bool) -> torch.Tensor:
    if not serialize:
        return packed.view(L, L).t()
    out = torch.zeros(L, L, dtype=torch.float64, device=packed.device)
    iu = torch.triu_indices(L, L, device=packed.device)
    # column-packed upper: (col i, row j<=i) at i(i+1)/2 + j  (structure.h:39)
    out[iu[0], iu[1]] = packed[(iu[1] * (iu[1] + 1)) // 2 + iu[0]]
    return out


def construct_R(args: info, topo=None) -> torch.Tensor:
    """rows x cols local block with zero lower part (serialize<uppertri, rect>, cholinv.hpp:30-37)."""
    return _expand(args.R, args.local_dim, args.serialize)


def construct_Rinv(args: info, topo=None) -> torch.Tensor:
    return _expand(args.Rinv, args.local_dim, args.serialize)


def residual(A: matrix, args: info, topo) -> float:
    """cholesky::validate<Alg>::residual (test/cholesky/validate.hpp:7-49)."""
    ctx = topo.context()
    r = C.c_double()
    ctx.check(_lib.lib().capital_cholinv_residual_f64(ctx.handle, A.data.data_ptr(), A.num_rows_global,
                                                      _lib.UPPERTRI_PACKED if args.serialize else _lib.RECT,
                                                      args.R.data_ptr(), C.byref(r)))
    return float(r.value)
