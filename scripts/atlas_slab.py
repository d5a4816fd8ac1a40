"""End-to-end brie-quant style fit on an atlas-shaped slab (BASELINE config C5: 1M cells, pseudotime
covariate + LRT) through the public API with HOST sparse layers: device filter statistics, sparse
ingest, chunked fit, memory-mapped dense outputs.  Prints a JSON line with the phase timings.

  python scripts/atlas_slab.py [cells] [events] [min_iter]      (defaults 1000000 256 300)
"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.sparse import csc_matrix
from brie_b200.models import fitBRIE
from brie_b200.utils.anndata_lite import AnnDataLite
from brie_b200.utils.preprocessing import filter_genes
from brie_b200.utils.synth import simulate_counts_device

Nc = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
Ng = int(sys.argv[2]) if len(sys.argv) > 2 else 256
min_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 300
t0 = time.time()
sim = simulate_counts_device(Nc, Ng, design='pseudotime', seed=5, with_efflen=True, n_layers=3, pseudo_count=0.0)
layers = {}
for key, t in zip(('isoform1', 'isoform2', 'ambiguous'), sim['layers']):
    layers[key] = csc_matrix(t[:, :Ng].cpu().numpy())          # what brie-count hands over (io_utils.py:107)
nnz = sum(v.nnz for v in layers.values())
del sim
torch.cuda.empty_cache()
t1 = time.time()
X = layers['isoform1'] + layers['isoform2'] + layers['ambiguous']
ad = AnnDataLite(X=X, layers=layers, varm={'effLen': np.tile(np.array([[172., 0, 284, 0, 72, 284]], np.float32), (Ng, 1))})
Xc = np.random.default_rng(0).uniform(0, 1, (Nc, 1)).astype(np.float32)
t2 = time.time()
filter_genes(ad, min_counts=50, min_counts_uniq=10, min_cells_uniq=30, min_MIF_uniq=0.001, device="cuda")
torch.cuda.synchronize()
t3 = time.time()
out_dir = tempfile.mkdtemp(prefix="brie_layers_")
res = fitBRIE(ad, Xc=Xc, LRT_index=None, intercept_mode='gene', min_iter=min_iter, max_iter=min_iter, MC_size=3,
              n_eval=500, out_dir=out_dir)
torch.cuda.synchronize()
t4 = time.time()
steps = int(min_iter / 6) * 6
print(json.dumps(dict(cells=Nc, events_in=Ng, events_kept=int(ad.shape[1]), stored_counts=int(nnz), models=2,
                      steps=steps, simulate_s=round(t1 - t0, 2), device_filter_s=round(t3 - t2, 2),
                      fit_total_s=round(t4 - t3, 2),
                      fit_value=Nc * ad.shape[1] * 3 * 2 * steps / (t4 - t3),
                      psi_memmap=isinstance(ad.layers['Psi'], np.memmap),
                      out_bytes=sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir)),
                      fdr05=int((res.fdr < 0.05).sum()))))
