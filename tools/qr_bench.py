"""CholeskyQR2 timing (BASELINE config 4: m=2^20, n=256, 1D grid) -- run plain (1 GPU) or under torch.distributed.run."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import capital_b200 as cb
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
m = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
topo = cb.topo.rect(world, rank, 1)
A = cb.matrix(n, m, 1, world).distribute_random(topo, rank)
args = cb.cacqr.info(2, cb.cholinv.info(0, 1, 0, "U"))
for _ in range(3):
    cb.cacqr.factor(A, args, topo)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1: dist.barrier()
torch.cuda.synchronize()
K = 10
e0.record()
for _ in range(K):
    cb.cacqr.factor(A, args, topo)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device="cuda")
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
res, orth = cb.cacqr.validate(A, args, topo)
if rank == 0:
    fl = 4 * m * n * n + 5 * n ** 3 / 3
    print(json.dumps({"alg": "cacqr2_1d", "m": m, "n": n, "n_gpus": world, "ms": t.item(), "tflops": fl / (t.item() * 1e-3) / 1e12,
                      "residual": res, "orthogonality": orth}))
if world > 1:
    dist.barrier(); cb.topo.release_contexts(); dist.destroy_process_group()
