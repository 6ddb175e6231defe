"""N>1 host logic on CPU (gloo): the SUMMA rank arithmetic used by dist.cu, run with numpy operands on the reference's
2x2x2 grid (8 processes) and the 1D CholeskyQR2 reduction on 2 processes, against the global-view oracle."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _summa_worker(rank, world, port, q):
    from capital_b200 import schedule as sch
    from oracle import capital_oracle as co
    _init(rank, world, port)
    c = d = 2
    x, y, z = sch.coords(c, d, rank)
    K, M, N = 12, 8, 10
    rng = np.random.default_rng(0)
    Xg, Yg = rng.standard_normal((K, M)), rng.standard_normal((K, N))
    Xl, Yl = co.cyclic_local(Xg, d, d, x, y), co.cyclic_local(Yg, d, d, x, y)
    plan = sch.summa_plan(c, d, rank)
    reqs = []
    for dst in plan["send_x_to"]:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(Xl)), dst, tag=1))
    for dst in plan["send_y_to"]:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(Yl)), dst, tag=2))
    Xu, Yu = torch.from_numpy(np.ascontiguousarray(Xl)).clone(), torch.from_numpy(np.ascontiguousarray(Yl)).clone()
    if plan["src_x"] != rank:
        dist.recv(Xu, plan["src_x"], tag=1)
    if plan["src_y"] != rank:
        dist.recv(Yu, plan["src_y"], tag=2)
    for r in reqs:
        r.wait()
    P = Xu.numpy().T @ Yu.numpy()  # local product on the k = z slice
    g = dist.new_group(plan["depth_group"]) if False else None
    # depth all-reduce: emulate with an all_gather over the world and a sum over my depth group
    allP = [torch.zeros_like(torch.from_numpy(P)) for _ in range(world)]
    dist.all_gather(allP, torch.from_numpy(P))
    C = sum(allP[r].numpy() for r in plan["depth_group"])
    ref = co.cyclic_local(Xg.T @ Yg, d, d, x, y)
    ok = np.allclose(C, ref, atol=1e-12)
    # global transpose = exchange with the partner + local transpose (util.hpp:232-247)
    T = torch.from_numpy(np.ascontiguousarray(Xl)).clone()
    p = plan["transpose_partner"]
    if p != rank:
        s = dist.isend(torch.from_numpy(np.ascontiguousarray(Xl)), p, tag=3)
        dist.recv(T, p, tag=3)
        s.wait()
    ok &= np.allclose(T.numpy().T, co.cyclic_local(Xg.T, d, d, x, y))
    # slice group ordering = slice rank x + d*y (block_to_cyclic expects it, util.hpp:56-96)
    ok &= plan["slice_group"].index(rank) == x + d * y
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def _general_worker(rank, world, port, q, c, d):
    """dist.cu::product() on the library's own non-cubic grids (2x1x1: k range split over the layers; 1x2x2: two SUMMA steps)."""
    from capital_b200 import schedule as sch
    from oracle import capital_oracle as co
    _init(rank, world, port)
    x, y, z = sch.coords(c, d, rank)
    K, M, N = 16, 8, 12
    rng = np.random.default_rng(1)
    Xg, Yg = rng.standard_normal((K, M)), rng.standard_normal((K, N))
    Xl, Yl = co.cyclic_local(Xg, d, d, x, y), co.cyclic_local(Yg, d, d, x, y)
    P = np.zeros((M // d, N // d))
    for sl in sch.product_slices(c, d, rank, K // d):
        r0, r1 = sl["rows"]
        xs, ys = np.ascontiguousarray(Xl[r0:r1]), np.ascontiguousarray(Yl[r0:r1])
        reqs = [dist.isend(torch.from_numpy(xs), dst, tag=10 + sl["kb"]) for dst in sl["send_x_to"]]
        reqs += [dist.isend(torch.from_numpy(ys), dst, tag=20 + sl["kb"]) for dst in sl["send_y_to"]]
        xu, yu = torch.from_numpy(xs.copy()), torch.from_numpy(ys.copy())
        if sl["src_x"] != rank:
            dist.recv(xu, sl["src_x"], tag=10 + sl["kb"])
        if sl["src_y"] != rank:
            dist.recv(yu, sl["src_y"], tag=20 + sl["kb"])
        for r in reqs:
            r.wait()
        P += xu.numpy().T @ yu.numpy()
    allP = [torch.zeros(P.shape, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allP, torch.from_numpy(P))
    C = sum(allP[sch.rank_of(c, d, x, y, zz)].numpy() for zz in range(c))  # depth all-reduce
    ok = np.allclose(C, co.cyclic_local(Xg.T @ Yg, d, d, x, y), atol=1e-12)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def _qr_worker(rank, world, port, q):
    from oracle import capital_oracle as co
    _init(rank, world, port)
    m, n = 64, 8
    blocks = [co.random_local(m, n, 1, world, 0, r, r) for r in range(world)]
    mine = blocks[rank]
    g = torch.from_numpy(np.triu(mine.T @ mine))
    dist.all_reduce(g)  # MPI_Allreduce of the Gram matrix (cacqr/policy.h:82)
    qs, r_or = co.cacqr_1d(blocks, 1)
    import scipy.linalg as sla
    G = g.numpy()
    r = sla.cholesky(G + np.triu(G, 1).T)
    ok = np.allclose(np.triu(r), r_or, atol=1e-12)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def _spawn(fn, world, port, *extra):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=fn, args=(r, world, port, q) + extra) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_summa_rank_arithmetic_on_2x2x2_gloo():
    _spawn(_summa_worker, 8, 29611)


def test_product_on_2x1x1_world_size_2_gloo():
    _spawn(_general_worker, 2, 29613, 2, 1)


def test_product_on_1x2x2_world_size_4_gloo():
    _spawn(_general_worker, 4, 29614, 1, 2)


def test_cacqr_1d_gram_allreduce_world_size_2_gloo():
    _spawn(_qr_worker, 2, 29612)


def test_plans_are_consistent():
    from capital_b200 import schedule as sch
    c = d = 2
    for r in range(8):
        p = sch.summa_plan(c, d, r)
        for dst in p["send_x_to"]:
            assert sch.summa_plan(c, d, dst)["src_x"] == r
        for dst in p["send_y_to"]:
            assert sch.summa_plan(c, d, dst)["src_y"] == r
        assert sch.summa_plan(c, d, p["transpose_partner"])["transpose_partner"] == r
