"""The north_star target run: a full brie-quant-style DAS fit with LRT of a BASELINE configuration on all GPUs of the
box, end to end through the public API (`fitBRIE`, what `brie-quant` calls after reading its input), from HOST sparse
(CSC) count layers to the output container + result table:

  torchrun --nproc-per-node 8 scripts/target_run.py C5 [--out-root /dev/shm] [--min-iter 5000] [--max-iter 20000]

  C5 : 1M cells x 20 000 events, pseudotime covariate + LRT (M = 2), events sharded over the ranks
  C4 : 200k cells x 20 000 genes, spliced/unspliced, Kg = 8 gene features, interceptMode cell (shared Wg: per-step
       NCCL all-reduce issued by the library)
  C3 : 100k cells x 10 000 events, 3 covariates + LRT (M = 4)

Every rank draws the counts of ITS event shard on the device (synthetic, SURVEY 8d recipe), brings them to the host as
scipy CSC (untimed; the other shards' columns stay empty -- a rank only ever reads its own), then the timed region is
fitBRIE: device ingest, pseudo-count, the full default schedule with per-batch convergence extensions, loss_gene,
batched LRT, p / FDR, D2H of the dense layers into .npy maps under --out-root (Psi, Psi_95CI, Z_std; skipped with a
note if the file system cannot hold them), and rank 0 writes the container + `.brie_ident.tsv`.
Prints one JSON line: wall times, per-rank phases, fused-kernel roofline fraction per rank (CUDA events around every
schedule launch), SM clocks / power under load.
"""
import argparse, json, os, shutil, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch

CONFIGS = {
    "C2": dict(cells=5000, events=5000, design='binary1', layers=3, eff=True, Kg=0, mode='gene', lrt=True),
    "C3": dict(cells=100000, events=10000, design='mixed3', layers=3, eff=True, Kg=0, mode='gene', lrt=True),
    "C4": dict(cells=200000, events=20000, design='none', layers=2, eff=False, Kg=8, mode='cell', lrt=False),
    "C5": dict(cells=1000000, events=20000, design='pseudotime', layers=3, eff=True, Kg=0, mode='gene', lrt=True),
}
KEYS = ('isoform1', 'isoform2', 'ambiguous')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("--cells", type=int)
    ap.add_argument("--events", type=int)
    ap.add_argument("--min-iter", type=int, default=5000)
    ap.add_argument("--max-iter", type=int, default=20000)
    ap.add_argument("--n-eval", type=int, default=500)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--out-root", default="/dev/shm")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    Nc, Ng = a.cells or cfg['cells'], a.events or cfg['events']
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    from scipy.sparse import csc_matrix, hstack
    from brie_b200.models import fitBRIE
    from brie_b200.sharding import event_shards
    from brie_b200.utils import io_utils
    from brie_b200.utils.anndata_lite import AnnDataLite
    from brie_b200.utils.synth import make_design, simulate_counts_device

    M = 1 + (make_design(4, cfg['design'], np.random.default_rng(0)).shape[1] if cfg['lrt'] else 0)
    n_layers = cfg['layers']
    group = max(1, int(np.ceil(500000 / Nc))) if (cfg['Kg'] == 0 and cfg['mode'] != 'cell') else 1
    lo, hi = event_shards(Ng, world, group)[rank]

    # ---- where do the dense output layers go?
    need = 3 * Nc * Ng * 4
    out_dir, layer_note = None, None
    if rank == 0:
        try:
            free = shutil.disk_usage(a.out_root).free
        except OSError:
            free = 0
        ok = free > need * 1.05 + (8 << 30)
        info = [ok, free]
    else:
        info = [None, None]
    if dist is not None:
        dist.broadcast_object_list(info, src=0)
    layers_ok, free = info
    out_dir = os.path.join(a.out_root, "brie_target_%s" % a.config)
    out_keys = ('Psi', 'Psi95CI', 'Z_std') if layers_ok else ('Psi',)
    if not layers_ok:
        layer_note = ("%s has %.0f GB free, the three dense float32 layers need %.0f GB: only Psi is kept if it fits, "
                      "else none" % (a.out_root, free / 1e9, need / 1e9))
        if free < need / 3 * 1.05 + (8 << 30):
            out_keys = ()

    # ---- synthetic input: this rank's shard, drawn on the device, held on the host as CSC (untimed)
    t0 = time.time()
    Xc = make_design(Nc, cfg['design'], np.random.default_rng(0))
    keys = KEYS[:n_layers] if cfg['eff'] else ('spliced', 'unspliced')
    blocks = {k: [csc_matrix((Nc, lo), dtype=np.float32)] for k in keys}
    eff = np.ones((Ng, 6), np.float32)
    n_counts = np.zeros(Ng); n_uniq = np.zeros(Ng); cdr = np.zeros(Ng)
    for b0 in range(lo // a.block * a.block, hi, a.block):       # global blocks: the data do not depend on the sharding
        nb = min(a.block, Ng - b0)
        sim = simulate_counts_device(Nc, nb, design=cfg['design'], seed=100 + b0, with_efflen=cfg['eff'],
                                     n_layers=n_layers, pseudo_count=0.0, event_offset=b0, Xc=Xc, device="cuda:%d" % local)
        s0, s1 = max(lo, b0) - b0, min(hi, b0 + nb) - b0
        e0, n = b0 + s0, s1 - s0
        lay = [t[:, s0:s1] for t in sim['layers']]
        tot = sum(lay)
        n_counts[e0:e0 + n] = tot.sum(0).double().cpu().numpy()
        n_uniq[e0:e0 + n] = (lay[0] + lay[1]).sum(0).double().cpu().numpy()
        cdr[e0:e0 + n] = (tot > 0).float().mean(0).cpu().numpy()
        for k, t in zip(keys, lay):
            blocks[k].append(bench.device_block_to_csc(t, n))
        if cfg['eff']:
            eff[e0:e0 + n] = sim['effLen'][s0:s1]
        del sim, tot, lay
    for k in keys:
        blocks[k].append(csc_matrix((Nc, Ng - hi), dtype=np.float32))
    layers = {k: hstack(v, format='csc') for k, v in blocks.items()}
    del blocks
    torch.cuda.empty_cache()
    nnz = int(sum(v.nnz for v in layers.values()))
    ad = AnnDataLite(X=layers[keys[0]], layers=layers, varm={'effLen': eff} if cfg['eff'] else {})
    Xg = np.random.default_rng(1).standard_normal((Ng, cfg['Kg'])).astype(np.float32) if cfg['Kg'] else None
    t_sim = time.time() - t0

    os.environ["BRIE_KERNEL_TIMING"] = str(int(a.min_iter / 6) * 6)      # every schedule launch of every chunk
    os.environ["BRIE_TIMING"] = "1"
    sampler = bench.ClockSampler(local, 500)
    sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.time()
    res = fitBRIE(ad, Xc=Xc if Xc.shape[1] else None, Xg=Xg, LRT_index=None if cfg['lrt'] else [],
                  intercept_mode=cfg['mode'], layer_keys=list(keys), min_iter=a.min_iter, max_iter=a.max_iter,
                  MC_size=3, n_eval=a.n_eval, seed=7, out_dir=out_dir, out_keys=out_keys)
    torch.cuda.synchronize()
    t_fit_local = time.time() - t1
    if dist is not None:
        dist.barrier()
    t2 = time.time()
    clocks = sampler.stop(t1, t2)

    # ---- what brie-quant does after the fit (quant.py:118-130): container + result table, rank 0
    stats = [n_counts[lo:hi], n_uniq[lo:hi], cdr[lo:hi]]
    if dist is not None:
        parts = [None] * world
        dist.all_gather_object(parts, stats)
        stats = [np.concatenate([p[i] for p in parts]) for i in range(3)]
    t_write = None
    if rank == 0:
        ad.var['n_counts'], ad.var['n_counts_uniq'], ad.var['cdr'] = stats[0], stats[1], stats[2]
        ad.uns['brie_version'] = "b200"
        ad.uns['Xc_ids'] = np.array(["x%d" % i for i in range(Xc.shape[1])])
        ad.uns['Xg_ids'] = None
        for k in keys:                          # the container of the synthetic input is not part of the measurement
            del ad.layers[k]
        ad.X = csc_matrix((Nc, Ng), dtype=np.float32)
        os.makedirs(out_dir, exist_ok=True)        # (a single-process fit with shared parameters keeps its layers in RAM)
        out_file = os.path.join(out_dir, "brie_quant.npz")
        ad.write_npz(out_file)
        df = io_utils.dump_results(ad)
        df.to_csv(out_file[:-4] + '.brie_ident.tsv', sep='\t', header=True, index=True, index_label='GeneID',
                  float_format='%.3e')
        t_write = time.time() - t2
    # ---- per-rank report
    tm = getattr(res, 'timing_rank', None) or getattr(res, 'timing', None) or {}    # this rank's own phases / kernel times
    peak = bench.hbm_peak()[0]
    frac = None
    if tm.get("step_kernel_ms_sum"):
        frac = tm["step_kernel_alg_bytes_sum"] / (tm["step_kernel_ms_sum"] * 1e-3) / 1e9 / peak
    mine = dict(rank=rank, events=[lo, hi], fit_s=round(t_fit_local, 2), simulate_s=round(t_sim, 1),
                phases={k: round(v, 2) for k, v in tm.items() if not k.startswith("step_kernel")},
                step_kernel_launches=tm.get("step_kernel_launches"), step_kernel_ms_mean=(
                    tm["step_kernel_ms_sum"] / tm["step_kernel_launches"] if tm.get("step_kernel_launches") else None),
                step_kernel_frac_of_hbm_peak=frac, clocks=clocks, stored_counts=nnz)
    allr = [mine]
    if dist is not None:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    if rank == 0:
        n_iter = np.asarray(res.n_iter)
        steps_min = int(a.min_iter / 6) * 6
        fr = [r["step_kernel_frac_of_hbm_peak"] for r in allr if r["step_kernel_frac_of_hbm_peak"]]
        out = dict(config=a.config, cells=Nc, events=Ng, n_gpus=world, models=M, mc_size=3, min_iter=a.min_iter,
                   max_iter=a.max_iter, host_input="scipy CSC layers (each rank holds its shard's columns)",
                   fit_lrt_wall_s=round(t2 - t1, 2), container_and_table_s=None if t_write is None else round(t_write, 2),
                   step_kernel_frac_of_hbm_peak_min=min(fr) if fr else None,
                   step_kernel_frac_of_hbm_peak_mean=float(np.mean(fr)) if fr else None,
                   hbm_peak_GBps=peak, n_iter_mean=float(n_iter.mean()), n_iter_max=int(n_iter.max()),
                   value_at_min_schedule=Nc * Ng * 3.0 * M * steps_min / (t2 - t1),
                   unit="cell*event*sample/s over the fit+LRT wall (schedule steps only in the numerator)",
                   output_layers=list(out_keys), output_dir=out_dir, layer_note=layer_note,
                   psi_mean_first_events=float(np.asarray(ad.layers['Psi'][:, :4]).mean()) if out_keys else None,
                   fdr05_calls=[int(v) for v in (np.asarray(res.fdr) < 0.05).sum(0)] if cfg['lrt'] else None,
                   ranks=allr)
        line = json.dumps(out)
        print(line)
        if a.json:
            open(a.json, "w").write(line + "\n")
    if dist is not None:
        dist.barrier()
        from brie_b200 import comm
        comm.destroy_all()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
