"""Consistency probe of the distributed cholinv::factor: residual, bit-identical replicas across the depth layers, run-to-run
determinism.  torchrun --nproc-per-node N tools/dist_check.py n bcm c reps   (CAPITAL_MP_SAME_DEVICE=1: the ranks share the visible
GPUs round-robin -- 8 ranks on 1 GPU, or 4 + 4 on two GPUs with real NVLink traffic between the halves)"""
import os, sys, hashlib
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import capital_b200 as cb


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n, bcm, c, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    same = bool(os.environ.get("CAPITAL_MP_SAME_DEVICE"))  # ranks share the visible GPUs round-robin (gloo bootstrap, CUDA IPC)
    torch.cuda.set_device(lr % torch.cuda.device_count() if same else lr)
    dist.init_process_group("gloo" if same else "nccl", **({} if same else {"device_id": torch.device("cuda", lr)}))
    topo = cb.topo.square(world, rank, c)
    A = cb.matrix(n, n, topo.d, topo.d).distribute_symmetric(topo)
    sums = []
    for rep in range(reps):
        args = cb.cholinv.info(0, 1, bcm, "U")
        cb.cholinv.factor(A, args, topo)
        res = cb.cholinv.residual(A, args, topo)
        h = hashlib.sha1(args.R.cpu().numpy().tobytes() + args.Rinv.cpu().numpy().tobytes()).hexdigest()[:12]
        allh = [None] * world
        dist.all_gather_object(allh, (topo.x, topo.y, topo.z, h, res))
        sums.append(h)
        if rank == 0:
            byxy = {}
            for x, y, z, hh, r in allh:
                byxy.setdefault((x, y), set()).add(hh)
            rep_ok = all(len(v) == 1 for v in byxy.values())
            print(f"n={n} bcm={bcm} rep {rep}: residual {max(a[4] for a in allh):.3e} replicas identical: {rep_ok} hashes {sorted(set(a[3] for a in allh))[:4]}", flush=True)
    if rank == 0:
        print(f"  deterministic across reps (rank 0): {len(set(sums)) == 1}; peer flag waits: {topo.context().peer_wait_mode()}", flush=True)
    dist.barrier()
    cb.topo.release_contexts()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
