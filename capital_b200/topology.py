"""Process grids of the reference (src/util/topology.h): `topo.square` (c x d x d) and `topo.rect` (c x d x c).
MPI sub-communicators are replaced by the library's peer layer (capital_comm_init*: every rank maps every other rank's work
arena over NVLink; NCCL or torch.distributed only exchanges the handles); one process per GPU, rendezvous through torch.distributed."""
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
        import torch
        key = tuple(getattr(self.grid, f) for f, _ in self.grid._fields_)
        ctx = _contexts.get(key)
        if ctx is None:
            if device is None:
                device = torch.cuda.current_device()
            # run on torch's current stream so that library calls are ordered with the caller's tensor ops
            # (handle 0 is the legacy default stream: pass cudaStreamLegacy = 0x1, since NULL means "library-owned")
            stream = torch.cuda.current_stream(device).cuda_stream or 0x1
            ctx = _lib.Context(self.grid, device, stream)
            ctx._device, ctx._stream = device, stream
            if self.size > 1:
                _comm_init(ctx, self)
            _contexts[key] = ctx
        else:
            # the caller may have switched torch streams since the context was made: follow it, so that library work stays
            # ordered with the producer ops of A / R / Rinv on the stream that is current NOW
            stream = torch.cuda.current_stream(ctx._device).cuda_stream or 0x1
            if stream != ctx._stream:
                ctx.set_stream(stream)
                ctx._stream = stream
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
    """Join the clique of the grid's ranks (capital_comm_init*): the library maps every rank's work arena into every other rank
    (CUDA IPC) and needs a host allgather for the handles.  Default: NCCL (rank 0's ncclUniqueId broadcast through
    torch.distributed).  With a gloo process group, or CAPITAL_BOOTSTRAP=host, torch.distributed itself is the allgather -- which
    also allows several ranks to share one GPU (tests), something NCCL refuses."""
    import os
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("multi-GPU grid: initialise torch.distributed first (one process per GPU)")
    backend = dist.get_backend()
    if backend == "gloo" or os.environ.get("CAPITAL_BOOTSTRAP") == "host":
        world = topo.size
        on_gpu = backend != "gloo"

        def allgather(user, send, recv, nbytes):
            try:
                mine = torch.frombuffer(C.string_at(send, nbytes), dtype=torch.uint8).clone()
                if on_gpu:
                    mine = mine.cuda()
                parts = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine)
                out = torch.cat(parts).cpu().numpy().tobytes()
                C.memmove(recv, out, nbytes * world)
                return 0
            except Exception:  # noqa: the C side reports the failure
                return 1

        ctx._allgather = _lib.ALLGATHER_FN(allgather)  # keep the trampoline alive as long as the context
        ctx.check(_lib.lib().capital_comm_init_host(ctx.handle, ctx._allgather, None))
        return
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
