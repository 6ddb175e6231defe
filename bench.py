#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: FP64 CholInv (cholesky::cholinv::factor) on synthetic SPD input.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n SIZE]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one cholinv::factor of the workload below on a resident (HBM) copy of the reference's own
generator matrix (structure.hpp:69-103).  Rank 0 prints ONE JSON line.  `value` = whole-job n^3/3 / time
("Cholesky TFLOP/s", the BASELINE.json metric); `e2e` repeats it through the public Python API with pinned HOST
buffers (H2D of A and D2H of R, Rinv inside the timed region); `roofline` times the dominant kernel (128x128 DMMA
GEMM) with CUDA events on its own stream inside the timed steps; `cpu_baseline` / `--impl reference` time the
reference's own CPU implementation (oracle/_ref, built from /root/reference by oracle/build_ref.sh) on this box's
host cores.  Workloads (BASELINE.json configs): N=1 n=16384 b=512 | N=8 n=65536 b=1024 on the reference's 2x2x2 grid;
N=2 / N=4 are not valid reference grids (summa.hpp:16-31 needs c == d) and run the library's own 2x1x1 / 1x2x2 grids.
"""
from __future__ import annotations
import argparse, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DMMA_PEAK_TFLOPS = 37.2  # measured on this pool's B200: profiles/r01_fp64_pipe_ceilings.log (DMMA.8x8x4, 148 SMs x 64 FMA/clk x 1.965 GHz)
WORKLOADS = {  # n_gpus -> (n, c, bc_mult_dim)  ; base-case size b = 512 (N=1) / 1024 (N=8) as in BASELINE.json configs
    1: (16384, 1, -5),
    2: (24576, 2, -5),
    4: (32768, 1, -5),
    8: (65536, 2, -4),
}


def clocks_sampler(stop: threading.Event, out: list, gpu_index: int):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 7:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples: list) -> dict:
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no nvidia-smi samples"]}
    sm = sorted(float(s[0]) for s in samples)
    reasons = []
    for idx, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
        if any(s[idx].lower().startswith("active") for s in samples):
            reasons.append(name)
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]), "power_w_max": max(float(s[2]) for s in samples),
            "samples": len(samples), "reasons": reasons}


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "ref_cholinv")
    return p if os.path.exists(p) else None


def run_reference_cholinv(n: int, bc_mult: int, ranks: int, iters: int, timeout: float):
    """Time the reference's own cholinv::factor on host cores (bench/cholesky/cholinv.cpp protocol). Returns dict or None."""
    exe = ref_binary()
    if exe is None:
        return None
    cores = os.cpu_count() or 1
    threads = max(1, cores // ranks)
    env = dict(os.environ, MINIMPI_NP=str(ranks), OPENBLAS_NUM_THREADS=str(threads), OMP_NUM_THREADS=str(threads))
    policy = 2 if ranks == 1 else 0  # benchmarked NoReplication policy is only valid at P == 1
    try:
        r = subprocess.run([exe, str(n), "0", "1", str(bc_mult), str(policy), str(iters)], env=env, capture_output=True,
                           text=True, timeout=timeout)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        d["threads_per_rank"] = threads
        d["cores"] = min(cores, threads * ranks)
        return d
    except Exception as e:  # noqa
        return {"error": repr(e)[:200]}


def reference_arm(args):
    """--impl reference: the reference's CPU path on this box's host cores, same metric/unit/config keys."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_full, c, bcm = WORKLOADS[args.gpus]
    if args.n:
        n_full = args.n
    ranks = 1 if args.gpus == 1 else 8
    n = min(n_full, 16384 if ranks == 1 else 8192)  # bounded sample: the reference allocates ~30x the matrix
    bcm_s = bcm
    t0 = time.time()
    iters = min(max(1, args.steps), 5)  # bounded: one reference factorization of the n=16384 sample takes ~10 s on 128 cores
    d = run_reference_cholinv(n, bcm_s, ranks, iters, timeout=1500)
    if not d or "time_mean_s" not in d:
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/ref_cholinv missing or failed: {d}"}))
        return 0
    t = d["time_mean_s"]
    val = n ** 3 / 3 / t / 1e12
    sample = (f"n={n} (full workload n={n_full}); {ranks} rank(s) x {d['threads_per_rank']} OpenBLAS threads; {iters} timed factorizations after "
              f"one warm-up (bench/cholesky/cholinv.cpp protocol); reference validator residual {d['residual']:.2e}")
    out = {
        "impl": "reference", "metric": "cholesky_tflops_fp64", "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus, n_full, c, bcm),
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": d["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(out))
    return 0


def workload_config(n_gpus, n, c, bcm):
    d = {1: 1, 2: 1, 4: 2, 8: 2}[n_gpus]
    L = n // d
    bc = (L // min(L, (c * d) << (-bcm))) * d
    return {"workload": f"cholinv::factor n={n} FP64 SPD (distribute_symmetric, diagonally dominant), complete_inv=0 split=1 "
                        f"bc_mult_dim={bcm} (base case {bc}), grid c={c} d={d} ({n_gpus} GPU)",
            "n": n, "grid": f"{c}x{d}x{d}", "base_case": bc, "l2": "inputs larger than L2 (no flush needed)"}


def ours(args):
    import torch
    import torch.distributed as dist
    import capital_b200 as cb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, c, bcm = WORKLOADS[args.gpus]
    if args.n:
        n = args.n
    topo = cb.topo.square(world, rank, c)
    ctx = topo.context()
    dgrid = topo.d

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    A = cb.matrix(n, n, dgrid, dgrid).distribute_symmetric(topo)
    pack = cb.cholinv.info(0, 1, bcm, "U")
    for _ in range(args.warmup):
        cb.cholinv.factor(A, pack, topo)
    # ---- timed region: resident inputs ----
    samples, stop = [], threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local_rank), daemon=True)
    barrier()
    if rank == 0:
        th.start()
    ctx.reset_counters()
    ctx.profile_begin()  # events around the dominant kernel's launches only (26 per step): negligible perturbation
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        cb.cholinv.factor(A, pack, topo)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    stop.set()
    k_ms, k_flops, k_launches = ctx.profile_end()
    cnt = ctx.counters()
    # The timed steps overlap kernels on two streams, which stretches event-bracketed launch durations.  The roofline
    # number is therefore taken from one extra step with the deferred stream disabled (same kernels, same launches).
    ctx.set_overlap(False)
    ctx.profile_begin()
    barrier()
    e0.record()
    cb.cholinv.factor(A, pack, topo)
    e1.record()
    barrier()
    s_ms_step = e0.elapsed_time(e1)
    s_ms, s_flops, s_launches = ctx.profile_end()
    ctx.set_overlap(True)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = n ** 3 / 3 / (ms_step * 1e-3) / 1e12
    residual = cb.cholinv.residual(A, pack, topo)

    # ---- end to end: pinned host buffers through the same public call ----
    e2e = None
    try:
        hostA = cb.matrix(n, n, dgrid, dgrid, data=A.data.cpu().pin_memory())
        hpack = cb.cholinv.info(0, 1, bcm, "U")
        cb.cholinv.factor(hostA, hpack, topo)  # warm-up: allocates pinned outputs + staging
        steps_e = max(1, min(args.steps, 3))
        ctx.reset_counters()
        barrier()
        e0.record()
        for _ in range(steps_e):
            cb.cholinv.factor(hostA, hpack, topo)
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        c2 = ctx.counters()
        ms_e = te.item() / steps_e
        e2e = {"value": n ** 3 / 3 / (ms_e * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms_e, "steps": steps_e,
               "h2d_bytes_per_step": c2.h2d_bytes // steps_e, "d2h_bytes_per_step": c2.d2h_bytes // steps_e,
               "note": "per rank; pinned host A in, pinned host R and Rinv (packed upper) out"}
        del hostA, hpack
    except Exception as ex:  # noqa
        e2e = {"value": None, "unit": "TFLOP/s", "error": repr(ex)[:200], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

    if rank == 0:
        ach = s_flops / (s_ms * 1e-3) / 1e12 if s_ms > 0 else None
        ach_ov = k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
        out = {
            "metric": "cholesky_tflops_fp64", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, n, c, bcm),
            "work_tflops": 5 * n ** 3 / 12 / (ms_step * 1e-3) / 1e12,  # CholInv with complete_inv=0 does 5n^3/12 flops
            "residual": residual,
            "roofline": {"bound": "tensor", "achieved": ach, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": (ach / DMMA_PEAK_TFLOPS) if ach else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the three launches captured with
                         # `ncu --set full` (profiles/r01e_gemm_tn_ncu_full.md: 0.06 / 0.68 / 8.53 GB for the 2048 / 4096 / 8192 levels)
                         "traffic": 3.09e9 if world == 1 else None,
                         "kernel": "gemm_tn_kernel<128,128,64,32,5> (DMMA.8x8x4 + TMA)", "launches": s_launches,
                         "kernel_share_of_step": s_ms / s_ms_step if s_ms_step else None,
                         "measured_in": "one extra step after the timed region with the deferred stream disabled (single-stream step "
                                        f"{s_ms_step:.2f} ms); inside the overlapped timed region the same launches read {ach_ov:.2f} TF/s "
                                        "because concurrent kernels share SMs" if ach_ov else None,
                         "peak_source": "measured DMMA pipe peak on this pool (profiles/r01_fp64_pipe_ceilings.log); MEASURED_PEAKS.json has no FP64 entry"},
            "e2e": e2e, "gpu_launches": int(cnt.kernel_launches), "clocks": summarize_clocks(samples),
        }
        if world == 1 and not args.no_cpu:
            d = run_reference_cholinv(min(n, 16384), bcm, 1, 1, timeout=600)
            if d and "time_mean_s" in d:
                ns = min(n, 16384)
                out["cpu_baseline"] = {"value": ns ** 3 / 3 / d["time_mean_s"] / 1e12, "unit": "TFLOP/s", "cores": d["cores"], "kind": "reference",
                                       "sample": f"reference cholinv::factor n={ns} same args, 1 rank x {d['threads_per_rank']} OpenBLAS threads, "
                                                 f"{d['time_mean_s']:.3f} s, residual {d['residual']:.2e}"}
            else:
                out["cpu_baseline"] = {"value": None, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {d}"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        cb.topo.release_contexts()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=0, help="override the matrix size (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.gpus not in WORKLOADS:
        raise SystemExit("--gpus must be one of 1, 2, 4, 8")
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    return reference_arm(args) if args.impl == "reference" else ours(args)


if __name__ == "__main__":
    sys.exit(main())
