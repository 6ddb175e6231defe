"""Developer timing loop (not the contract bench): cholinv::factor on resident data, a few sizes."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import capital_b200 as cb
topo = cb.topo.square(1, 0, 1)
ctx = topo.context()
sizes = [int(s) for s in sys.argv[1:]] or [2048, 4096, 8192, 16384]
for n in sizes:
    A = cb.matrix(n, n, 1, 1).distribute_symmetric(topo)
    bcm = 0
    while (n >> (-bcm)) > 512: bcm -= 1
    args = cb.cholinv.info(0, 1, bcm, "U")
    cb.cholinv.factor(A, args, topo)
    best = 1e9
    for _ in range(3):
        ctx.reset_counters()
        cb.cholinv.factor(A, args, topo)
        best = min(best, ctx.last_factor_ms())
    c = ctx.counters()
    res = cb.cholinv.residual(A, args, topo)
    print(f"n={n} bc_mult={bcm} {best:.2f} ms  cholesky(n^3/3)={n**3/3/best/1e9:.2f} TF/s  work(5n^3/12)={5*n**3/12/best/1e9:.2f} TF/s "
          f"launches={c.kernel_launches} gemms={c.gemm_launches} leaves={c.leaf_launches} residual={res:.2e}", flush=True)
    del A, args
