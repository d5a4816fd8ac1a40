"""Step-kernel throughput on the other BASELINE shapes (device-generated counts).

  C3 slab : 100k cells x 2048 events, Kc=3 (+LRT on all 3 -> M=4), 3 layers + effLen, gene intercept
  C4 slab : 200k cells x 4096 genes, spliced/unspliced (2 layers, no effLen), Kg=8, interceptMode cell, M=1
  C5 slab : 1M cells x 512 events, pseudotime covariate + LRT (M=2)
  W16     : 20k cells x 4096 genes, 15 covariates (detection rate + 14 cluster indicators, the dentate-gyrus
            design) in one model -> the 2-events-per-lane instantiation (Kc 15 -> 16)
Prints one JSON line per shape: ms/step, cell*event*sample/s, algorithmic GB/s and fraction of measured HBM.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from brie_b200.engine import FitEngine
from brie_b200.utils.synth import simulate_counts_device

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6553.3

SHAPES = {
    "C2": dict(Nc=5000, Ng=5000, design='binary1', eff=True, layers=3, masks=[[0], []], mode='gene', Kg=0),
    "C3": dict(Nc=100000, Ng=2048, design='mixed3', eff=True, layers=3,
               masks=[[0, 1, 2], [1, 2], [0, 2], [0, 1]], mode='gene', Kg=0),
    "C4": dict(Nc=200000, Ng=4096, design='none', eff=False, layers=2, masks=[[]], mode='cell', Kg=8),
    "C3F": dict(Nc=100000, Ng=10000, design='mixed3', eff=True, layers=3,
                masks=[[0, 1, 2], [1, 2], [0, 2], [0, 1]], mode='gene', Kg=0),      # the whole C3 on one GPU (108 GB)
    "C3a": dict(Nc=100000, Ng=2500, design='mixed3', eff=True, layers=3,
                masks=[[0, 1, 2], [1, 2], [0, 2], [0, 1]], mode='gene', Kg=0),      # ld 2528: rows 128-byte aligned only
    "C3b": dict(Nc=100000, Ng=2560, design='mixed3', eff=True, layers=3,
                masks=[[0, 1, 2], [1, 2], [0, 2], [0, 1]], mode='gene', Kg=0),      # ld 2560: 20 whole tiles
    "C3H": dict(Nc=100000, Ng=5000, design='mixed3', eff=True, layers=3,
                masks=[[0, 1, 2], [1, 2], [0, 2], [0, 1]], mode='gene', Kg=0),
    "W16": dict(Nc=20000, Ng=4096, design='wide15', eff=False, layers=2, masks=[list(range(15))], mode='gene', Kg=0),
    "C5": dict(Nc=1000000, Ng=512, design='pseudotime', eff=True, layers=3, masks=[[0], []], mode='gene', Kg=0),
    # wide designs (contractions as GEMMs around the fused kernel): 24 cell covariates; 20 gene features + cell intercept
    "W24": dict(Nc=20000, Ng=4096, design='wide15', eff=False, layers=2, masks=[list(range(24))], mode='gene', Kg=0, Kc_extra=9),
    "G20": dict(Nc=200000, Ng=2048, design='none', eff=False, layers=2, masks=[[]], mode='cell', Kg=20),
    "K8G8": dict(Nc=50000, Ng=4096, design='wide15', eff=False, layers=2, masks=[list(range(8))], mode='gene', Kg=8),
}


def run(name, steps=30, warm=5, loss=False):
    c = SHAPES[name]
    sim = simulate_counts_device(c['Nc'], c['Ng'], design=c['design'], seed=3, with_efflen=c['eff'], n_layers=c['layers'])
    nz = float(((sim['layers'][0] + sim['layers'][1] + (sim['layers'][2] if c['layers'] > 2 else 0)) > 0).float().mean())
    Xg = np.random.default_rng(0).standard_normal((c['Ng'], c['Kg'])).astype(np.float32) if c['Kg'] else None
    if c.get('Kc_extra'):                     # more covariates than the synthetic design draws: standard-normal extras
        sim['Xc'] = np.concatenate([sim['Xc'], np.random.default_rng(1).standard_normal(
            (c['Nc'], c['Kc_extra'])).astype(np.float32)], axis=1)
    eng = FitEngine(sim['layers'], effLen=sim['effLen'], Xc=sim['Xc'], Xg=Xg, masks=c['masks'],
                    intercept_mode=c['mode'], MC_size=3, seed=1, n_events=c['Ng'], trace_cap=8,
                    group_size=max(1, -(-500000 // c['Nc'])))
    eng.init_params()
    eng.begin_stage(0.01)
    eng.run_steps(warm)
    torch.cuda.synchronize()
    eng.kernel_timing(steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if loss:
        for i in range(steps):
            eng.run_steps(1, 0)
    else:
        eng.run_steps(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kms, kn = eng.kernel_time_ms()
    M = len(c['masks'])
    alg = c['Nc'] * c['Ng'] * (4 * c['layers'] + 48 * M)
    step_alg = alg + (c['Nc'] * c['Ng'] * 16 * M if eng.wide else 0)      # wide: + prior-mean / residual planes, written and read
    out = dict(shape=name, wide=bool(eng.wide), step_frac_of_measured_hbm=round(step_alg / (ms * 1e-3) / 1e9 / PEAK, 4), cells=c['Nc'], events=c['Ng'], models=M, Kc=eng.Kc_real, Kg=eng.Kg_real, mode=c['mode'],
               layers=c['layers'], loss_trace=loss, nonzero_fraction=round(nz, 4), ms_per_step=round(ms, 4),
               kernel_ms=round(kms / kn, 4), value=c['Nc'] * c['Ng'] * 3 * M / (ms * 1e-3),
               alg_GBps=round(alg / (kms / kn * 1e-3) / 1e9, 1), frac_of_measured_hbm=round(alg / (kms / kn * 1e-3) / 1e9 / PEAK, 4),
               state_GB=round(torch.cuda.memory_allocated() / 1e9, 1), rows_per_cta=eng.sizes.rows_per_cta)
    print(json.dumps(out), flush=True)
    del eng, sim
    torch.cuda.empty_cache()


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    only = [a for a in sys.argv[1:] if a.startswith("--")]        # --loss / --noloss: one variant only (ncu runs)
    for n in args or ["C2", "C3", "C4", "C5"]:
        if "--loss" not in only:
            run(n)
        if "--noloss" not in only:
            run(n, loss=True)
