"""Process grids of the reference (src/util/topology.h): `topo.square` (c x d x d) and `topo.rect` (c x d x c).
MPI sub-communicators are replaced by NCCL communicators created inside the library (capital_comm_init);
one process per GPU, rendezvous through torch.distributed."""
from __future__ import annotations
import ctypes as C
from . import _lib

_contexts: dict = {}


class _Topo:
    def __init__(self, grid: _lib.Grid):
        self.grid = grid
        for f, _ in grid._fields_:
            setattr(self, f, getattr(grid, f))

    def context(self, device: int | None = None) -> _lib.Context:
        """Library context bound to this grid (cached per process)."""
        key = tuple(getattr(self.grid, f) for f, _ in self.grid._fields_)
        ctx = _contexts.get(key)
        if ctx is None:
            import torch
            if device is None:
                device = torch.cuda.current_device()
            # run on torch's current stream so that library calls are ordered with the caller's tensor ops
            # (handle 0 is the legacy default stream: pass cudaStreamLegacy = 0x1, since NULL means "library-owned")
            stream = torch.cuda.current_stream(device).cuda_stream or 0x1
            ctx = _lib.Context(self.grid, device, stream)
            if self.size > 1:
                _comm_init(ctx, self)
            _contexts[key] = ctx
        return ctx


class square(_Topo):
    """topo::square(comm, c, layout, num_chunks) -- topology.h:67-143 (layout 0)."""

    def __init__(self, size: int, rank: int, c: int, layout: int = 0, num_chunks: int = 0):
        g = _lib.Grid()
        st = _lib.lib().capital_grid_square(size, rank, c, layout, num_chunks, C.byref(g))
        if st != _lib.OK:
            raise _lib.CapitalError(st, f"invalid square grid: size={size} c={c} (needs size == c*d*d, layout 0)")
        super().__init__(g)


class rect(_Topo):
    """topo::rect(comm, c, layout, num_chunks) -- topology.h:16-65."""

    def __init__(self, size: int, rank: int, c: int, layout: int = 0, num_chunks: int = 0):
        g = _lib.Grid()
        st = _lib.lib().capital_grid_rect(size, rank, c, layout, num_chunks, C.byref(g))
        if st != _lib.OK:
            raise _lib.CapitalError(st, f"invalid rect grid: size={size} c={c} (needs c*c | size, layout 0)")
        super().__init__(g)


def _comm_init(ctx: _lib.Context, topo: _Topo):
    """Broadcast rank 0's ncclUniqueId through torch.distributed and join the clique."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("multi-GPU grid: initialise torch.distributed first (one process per GPU)")
    buf = C.create_string_buffer(128)
    if topo.rank == 0:
        st = _lib.lib().capital_comm_unique_id(buf)
        if st != _lib.OK:
            raise _lib.CapitalError(st, "capital_comm_unique_id failed (libnccl.so.2 not loadable)")
    box = [bytes(buf.raw)]
    dist.broadcast_object_list(box, src=0)
    uid = C.create_string_buffer(box[0], 128)
    ctx.check(_lib.lib().capital_comm_init(ctx.handle, uid))


def release_contexts():
    for ctx in _contexts.values():
        ctx.close()
    _contexts.clear()
