import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CAPITAL_BC_DEBUG"] = "1"
import capital_b200 as cb
from capital_b200 import _lib
from oracle import capital_oracle as co
ctx = cb.topo.square(1, 0, 1).context()
for n in (512, 512, 256, 128):
    a = torch.from_numpy(co.spd_global(n)).cuda()
    R = torch.zeros(n * n, dtype=torch.float64, device="cuda"); Ri = torch.zeros_like(R)
    ctx.check(_lib.lib().capital_lapack_potrf_trtri_f64(ctx.handle, n, a.data_ptr(), n, R.data_ptr(), n, Ri.data_ptr(), n))
