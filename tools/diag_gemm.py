import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import capital_b200 as cb
from capital_b200 import _lib
ctx = cb.topo.square(1, 0, 1).context()
def run(m, n, k, lda, ldb, ldc, flags=0, alpha=1.0, beta=0.0):
    g = torch.Generator(device="cuda").manual_seed(1)
    fa = torch.rand(lda * m, dtype=torch.float64, device="cuda", generator=g) - 0.5
    fb = torch.rand(ldb * n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    fc = torch.rand(ldc * n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    A = fa.view(m, lda).t()[:k]; B = fb.view(n, ldb).t()[:k]; Cm = fc.view(n, ldc).t()[:m]
    ref = alpha * (A.t() @ B) + beta * Cm
    st = _lib.lib().capital_blas_gemm_tn_f64(ctx.handle, m, n, k, alpha, fa.data_ptr(), lda, fb.data_ptr(), ldb, beta, fc.data_ptr(), ldc, flags)
    ctx.synchronize()
    e = (Cm - ref).abs()
    bad = (e > 1e-10).nonzero()
    msg = f"m={m} n={n} k={k} lda={lda} ldb={ldb} ldc={ldc} st={st} maxerr={e.max().item():.3e} nbad={bad.shape[0]}"
    if bad.shape[0]:
        r, c = bad[:, 0], bad[:, 1]
        msg += f" rows[{r.min().item()},{r.max().item()}] cols[{c.min().item()},{c.max().item()}] rowmod128={sorted(set((r % 128).tolist()))[:12]} colmod128={sorted(set((c % 128).tolist()))[:12]}"
    print(msg, flush=True)
for (m, n, k) in [(256, 256, 64), (256, 256, 96), (256, 256, 112), (256, 256, 128), (256, 256, 1024), (2048, 1536, 1024)]:
    for pad in (0, 2, 6, 16):
        run(m, n, k, k + pad, k + pad, m)
run(2048, 1536, 1024, 1024, 1024, 2051)
run(2048, 1536, 1024, 1040, 1024, 2048)
run(2048, 1536, 1024, 1024, 1040, 2048)
run(2048, 1536, 1024, 1056, 1056, 2048)
run(2048, 1536, 1024, 1152, 1152, 2048)
