"""Phase timing of a full C2 fit (diagnostic)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from brie_b200.engine import FitEngine, LEARNING_RATES

layers, effLen, Xc = bench.make_c2(1)
idx = layers[0] + layers[1] > 0
for i in range(2):
    layers[i][idx] += np.float32(0.01)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
t0 = T()
eng = FitEngine(layers, effLen=effLen, Xc=Xc, masks=[[0], []], model_ids=[0, 1], MC_size=3, seed=7,
                group_size=100, trace_cap=833)
t1 = T(); print("engine build + H2D %.3f s" % (t1 - t0))
eng.init_params(); t2 = T(); print("init %.3f" % (t2 - t1))
for i, lr in enumerate(LEARNING_RATES):
    eng.begin_stage(lr); eng.run_steps(833, 0 if i == 5 else -1)
    t3 = T(); print("stage %d: %.3f s  (%.4f ms/step)" % (i, t3 - t2, (t3 - t2) / 833 * 1e3)); t2 = t3
tr = eng.group_trace(833); t4 = T(); print("group_trace %.3f" % (t4 - t3))
for frac in (1.0, 0.5, 0.1):
    act = np.zeros((2, 50), bool); act[:, :int(50 * frac)] = True
    eng.set_active_groups(act); t5 = T()
    eng.run_steps(500, 0); t6 = T(); print("extension active %.2f: %.3f s (%.4f ms/step)" % (frac, t6 - t5, (t6 - t5) / 500 * 1e3))
eng.set_active_groups(np.ones((2, 50), bool))
t7 = T(); lg = eng.eval_loss_gene(500); t8 = T(); print("eval_loss_gene(500) %.3f s" % (t8 - t7))
post = [t.cpu().numpy() for t in eng.posterior(0)]; z = eng.Z_loc[0, :, :5000].cpu().numpy(); t9 = T()
print("posterior + D2H %.3f s" % (t9 - t8))
