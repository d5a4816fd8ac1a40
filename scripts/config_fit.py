"""End-to-end fit + LRT of a BASELINE.json configuration (or the per-GPU share of one) through the
public API, `brie_b200.models.fitBRIE`, from HOST sparse layers (CSC, what brie-count hands over,
io_utils.py:107): device ingest, pseudo-count, chunked fit with the full optimisation schedule and
convergence extensions, loss_gene, batched LRT, p / FDR, dense outputs (RAM or .npy memory maps).
Prints one JSON line: the BASELINE metric "brie-quant fit+LRT wall time" and the effective
cell x event x sample rate of the whole call.

  python scripts/config_fit.py C3 [--events N] [--cells N] [--min-iter 5000] [--max-iter 20000]
                                  [--n-eval 500] [--out-dir DIR] [--block 512]

  C2 : 5k x 5k, 1 binary covariate + LRT (M = 2)            C3 : 100k x 10k, 3 covariates + LRT (M = 4)
  C4 : 200k x 20k spliced/unspliced, Kg = 8, interceptMode cell (no LRT)
  C5 : 1M x 20k, pseudotime + LRT (M = 2); per-GPU share of the 8-GPU run = --events 2500
Under torchrun the events are sharded over the ranks (fitBRIE does that itself).
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.sparse import csc_matrix, hstack

CONFIGS = {
    "C2": dict(cells=5000, events=5000, design='binary1', layers=3, eff=True, Kg=0, mode='gene', lrt=True),
    "C3": dict(cells=100000, events=10000, design='mixed3', layers=3, eff=True, Kg=0, mode='gene', lrt=True),
    "C4": dict(cells=200000, events=20000, design='none', layers=2, eff=False, Kg=8, mode='cell', lrt=False),
    "C5": dict(cells=1000000, events=20000, design='pseudotime', layers=3, eff=True, Kg=0, mode='gene', lrt=True),
}
KEYS = ('isoform1', 'isoform2', 'ambiguous')


def device_block_to_csc(t, n):
    """(Nc, ld) device tile -> host CSC of its first n columns (column-major non-zeros)."""
    tt = t[:, :n].t().contiguous()
    nz = tt != 0
    colrow = nz.nonzero()
    indptr = np.zeros(n + 1, np.int64)
    indptr[1:] = np.cumsum(nz.sum(1).cpu().numpy())
    return csc_matrix((tt[nz].cpu().numpy(), colrow[:, 1].to(torch.int32).cpu().numpy(), indptr),
                      shape=(t.shape[0], n))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("--cells", type=int)
    ap.add_argument("--events", type=int)
    ap.add_argument("--min-iter", type=int, default=5000)
    ap.add_argument("--max-iter", type=int, default=20000)
    ap.add_argument("--n-eval", type=int, default=500)
    ap.add_argument("--block", type=int, default=512)
    ap.add_argument("--out-dir")
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    Nc, Ng = a.cells or cfg['cells'], a.events or cfg['events']

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    from brie_b200.models import fitBRIE
    from brie_b200.utils.anndata_lite import AnnDataLite
    from brie_b200.utils.synth import make_design, simulate_counts_device

    t0 = time.time()
    Xc = make_design(Nc, cfg['design'], np.random.default_rng(0))
    blocks = {k: [] for k in KEYS[:cfg['layers']]}
    eff = []
    for b, e0 in enumerate(range(0, Ng, a.block)):            # synthetic counts: drawn on the device, kept sparse on the host
        n = min(a.block, Ng - e0)
        sim = simulate_counts_device(Nc, n, design=cfg['design'], seed=100 + b, with_efflen=cfg['eff'],
                                     n_layers=cfg['layers'], pseudo_count=0.0, event_offset=e0, Xc=Xc)
        for k, t in zip(blocks, sim['layers']):
            blocks[k].append(device_block_to_csc(t, n))
        if cfg['eff']:
            eff.append(sim['effLen'])
        del sim
    layers = {k: hstack(v, format='csc') for k, v in blocks.items()}
    del blocks
    torch.cuda.empty_cache()
    nnz = int(sum(v.nnz for v in layers.values()))
    X = sum(layers.values())
    nonzero_fraction = X.nnz / float(Nc * Ng)
    varm = {'effLen': np.concatenate(eff, 0)} if cfg['eff'] else {}
    ad = AnnDataLite(X=X, layers=layers, varm=varm)
    Xg = np.random.default_rng(1).standard_normal((Ng, cfg['Kg'])).astype(np.float32) if cfg['Kg'] else None
    t1 = time.time()

    torch.cuda.synchronize()
    t2 = time.time()
    res = fitBRIE(ad, Xc=Xc if Xc.shape[1] else None, Xg=Xg, LRT_index=None if cfg['lrt'] else [],
                  intercept_mode=cfg['mode'], layer_keys=list(layers), min_iter=a.min_iter, max_iter=a.max_iter,
                  MC_size=3, n_eval=a.n_eval, **({'out_dir': a.out_dir} if a.out_dir else {}))
    torch.cuda.synchronize()
    t3 = time.time()
    if rank != 0:
        return
    M = 1 + (Xc.shape[1] if cfg['lrt'] else 0)
    n_iter = np.asarray(res.n_iter) if hasattr(res, 'n_iter') else None
    steps_min = int(a.min_iter / 6) * 6
    out = dict(config=a.config, cells=Nc, events=Ng, n_gpus=world, models=M, mc_size=3, min_iter=a.min_iter,
               max_iter=a.max_iter, nonzero_fraction=round(nonzero_fraction, 4), stored_counts=nnz,
               host_input="scipy CSC layers", simulate_s=round(t1 - t0, 1),
               fit_lrt_wall_s=round(t3 - t2, 2),
               value_at_min_schedule=Nc * Ng * 3.0 * M * steps_min / (t3 - t2),
               unit="cell*event*sample/s (schedule steps only; extensions, loss_gene, ingest and D2H are inside the wall time)",
               psi_is_memmap=isinstance(ad.layers['Psi'], np.memmap),
               psi_mean=float(np.asarray(ad.layers['Psi'][:, :min(Ng, 64)]).mean()))
    if getattr(res, 'timing', None):
        out['phase_s'] = {k: round(v, 2) for k, v in res.timing.items()}
    if n_iter is not None:
        out['n_iter_mean'] = float(n_iter.mean())
        out['n_iter_max'] = int(n_iter.max())
    if cfg['lrt']:
        out['fdr05_calls'] = [int(v) for v in (res.fdr < 0.05).sum(0)]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
