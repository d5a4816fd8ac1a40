"""Does the step kernel slow down in a long run (power cap -> lower SM clock)?  Blocks of steps on one shape with
nvidia-smi clocks / power sampled per block.

  python scripts/sustained_probe.py C3 [n_blocks] [steps_per_block]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scale_shapes import SHAPES, PEAK
from brie_b200.engine import FitEngine
from brie_b200.utils.synth import simulate_counts_device

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
n_blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 6
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
c = SHAPES[name]
sim = simulate_counts_device(c['Nc'], c['Ng'], design=c['design'], seed=3, with_efflen=c['eff'], n_layers=c['layers'])
Xg = np.random.default_rng(0).standard_normal((c['Ng'], c['Kg'])).astype(np.float32) if c['Kg'] else None
eng = FitEngine(sim['layers'], effLen=sim['effLen'], Xc=sim['Xc'], Xg=Xg, masks=c['masks'], intercept_mode=c['mode'],
                MC_size=3, seed=1, n_events=c['Ng'], trace_cap=8, group_size=max(1, -(-500000 // c['Nc'])))
eng.init_params()
eng.begin_stage(0.01)
eng.run_steps(5)
torch.cuda.synchronize()
M = len(c['masks'])
alg = c['Nc'] * c['Ng'] * (4 * c['layers'] + 48 * M)
for b in range(n_blocks):
    s = bench.ClockSampler(0, 100)
    s.start()
    time.sleep(0.15)
    eng.kernel_timing(steps)
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run_steps(steps); e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    kms, kn = eng.kernel_time_ms()
    ck = s.stop(t0, t1)
    print(json.dumps(dict(shape=name, block=b, steps=steps, ms_per_step=round(e0.elapsed_time(e1) / steps, 4),
                          kernel_ms=round(kms / kn, 4), frac=round(alg / (kms / kn * 1e-3) / 1e9 / PEAK, 4),
                          sm_mhz=ck['sm_mhz'], power_w_max=ck.get('power_w_max'), reasons=ck['reasons'])), flush=True)
