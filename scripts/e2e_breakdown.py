"""Phase timing of the public-API C2 fit (bench.py's e2e leg) and of its engine internals (diagnostic)."""
import contextlib, io, os, sys, time
os.environ["BRIE_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from brie_b200.engine import FitEngine
from brie_b200.models import fit_BRIE_matrix

orig_fit = FitEngine.fit


def timed_fit(self, *a, **k):
    def T():
        torch.cuda.synchronize(); return time.perf_counter()
    ev, rs = self.eval_loss_gene, self.run_steps
    acc = {"eval": 0.0, "steps": 0.0, "nsteps": 0, "loss_steps": 0}

    def ev2(*aa, **kk):
        t = T(); r = ev(*aa, **kk); acc["eval"] += T() - t; return r

    def rs2(n, slot=-1):
        t = T(); r = rs(n, slot); acc["steps"] += T() - t; acc["nsteps"] += n
        acc["loss_steps"] += n if slot >= 0 else 0
        return r
    self.eval_loss_gene, self.run_steps = ev2, rs2
    t0 = T(); out = orig_fit(self, *a, **k); tot = T() - t0
    print("engine.fit %.3f s: steps %.3f s (%d launches, %d with loss trace), loss_gene eval %.3f s, host logic %.3f s"
          % (tot, acc["steps"], acc["nsteps"], acc["loss_steps"], acc["eval"], tot - acc["steps"] - acc["eval"]))
    return out


FitEngine.fit = timed_fit
layers, effLen, Xc = bench.make_c2(1)
torch.cuda.synchronize()
t0 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()) as buf:
    res = fit_BRIE_matrix(layers, Xc=Xc, effLen=effLen, intercept=None, intercept_mode='gene', LRT_index=None,
                          min_iter=5000, max_iter=20000, MC_size=3, group_size=100, seed=7)
torch.cuda.synchronize()
print(buf.getvalue().strip().splitlines()[-1])
print("fit_BRIE_matrix total %.3f s" % (time.perf_counter() - t0), res.timing)
