"""One warm-up + `reps` cholinv::factor calls on resident data (profiling target for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import capital_b200 as cb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bcm = 0
while (n >> (-bcm)) > 512: bcm -= 1
topo = cb.topo.square(1, 0, 1)
A = cb.matrix(n, n, 1, 1).distribute_symmetric(topo)
args = cb.cholinv.info(0, 1, bcm, "U")
for _ in range(1 + reps):
    cb.cholinv.factor(A, args, topo)
torch.cuda.synchronize()
print("done", topo.context().last_factor_ms(), "ms")
