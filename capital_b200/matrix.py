"""Local storage of a cyclically distributed matrix (src/matrix/matrix.h:9-97, structure.h:8-52): FP64,
column-major local block of ceil(global/grid) rows and columns, held in a torch tensor (HBM or pinned host).
Naming follows the reference: X = columns, Y = rows; constructor order (globalCols, globalRows, gridCols, gridRows)."""
from __future__ import annotations
import torch
from . import _lib


def local_dim(glob: int, grid: int) -> int:
    return glob // grid + (1 if glob % grid else 0)  # matrix.hpp:8-11


class matrix:
    def __init__(self, global_cols: int, global_rows: int, grid_cols: int, grid_rows: int, device="cuda", data=None):
        self.num_columns_global, self.num_rows_global = global_cols, global_rows
        self.num_columns_local, self.num_rows_local = local_dim(global_cols, grid_cols), local_dim(global_rows, grid_rows)
        n = self.num_columns_local * self.num_rows_local
        self.data = data if data is not None else torch.zeros(n, dtype=torch.float64, device=device)
        assert self.data.dtype == torch.float64 and self.data.numel() == n and self.data.is_contiguous()

    def num_elems(self) -> int:
        return self.data.numel()

    def view2d(self) -> torch.Tensor:
        """rows x cols view of the column-major block."""
        return self.data.view(self.num_columns_local, self.num_rows_local).t()

    def distribute_symmetric(self, topo, diagonally_dominant: bool = True):
        """matrix::distribute_symmetric(x, y, d, d, key, dd) -- structure.hpp:69-103 (key is irrelevant there)."""
        assert self.num_columns_global == self.num_rows_global
        ctx = topo.context()
        ctx.check(_lib.lib().capital_distribute_symmetric_f64(ctx.handle, self.data.data_ptr(), self.num_rows_global,
                                                              int(diagonally_dominant)))
        return self

    def distribute_random(self, topo, key: int):
        """matrix::distribute_random(x, y, c, d, key) -- structure.hpp:106-129."""
        ctx = topo.context()
        ctx.check(_lib.lib().capital_distribute_random_f64(ctx.handle, self.data.data_ptr(), self.num_rows_global,
                                                           self.num_columns_global, key))
        return self
