"""N>1 host logic on CPU with a real multi-process rendezvous (gloo): every rank dry-runs the distributed cholinv::factor schedule
of capital_b200/csrc/dist.cu for ITS OWN grid coordinates (capital_dist_trace_cholinv -- the real schedule code, no device), the
traces are all-gathered, and every rank replays the whole grid: the flag protocol must drain (no deadlock) and be race free.
World size 2 = the library's 2 x 1 x 1 grid, world size 4 = 1 x 2 x 2 (the reference's own 2 x 2 x 2 grid is replayed in-process by
test_dist_protocol.py)."""
import os, sys
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, c, d, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CAPITAL_DIST_FAR_MIN="64", CAPITAL_DIST_SIDE_MIN="32")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import test_dist_protocol as tp
    mine = tp.trace(world, rank, c, n, 1, -2)
    traces = [None] * world
    dist.all_gather_object(traces, mine)
    rp = tp.Replay(traces)
    stuck = rp.run((c, d))
    races = rp.races() if not stuck else []
    q.put((rank, len(stuck), len(races), len(mine)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,c,d,port", [(2, 2, 1, 29611), (4, 1, 2, 29613)])
def test_every_rank_agrees_the_protocol_drains(world, c, d, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, c, d, 1024, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] == 0 and r[2] == 0 and r[3] > 0 for r in res), res
