"""`simulator`: re-draw the read counts of a fitted AnnData given its observed depth -- drop-in for
brie.models.simulator (brie/models/simulator.py:7-75), the multinomial sampling on the device
(`brie_resample_counts`) instead of `tfd.Multinomial(...).sample()` on a dense (Nc, Ng, 3) tensor."""
import ctypes as C
import sys

import numpy as np
from scipy.special import expit


def _dense(x):
    from scipy.sparse import issparse
    return np.asarray(x.toarray() if issparse(x) else x, np.float32)


def simulator(adata, Psi=None, effLen=None, mode="posterior",
              layer_keys=['isoform1', 'isoform2', 'ambiguous'], prior_sigma=None, seed=None, device=None):
    """Simulate read counts for BRIE model.

    Psi : (cells, events) array to simulate from; default: mode "posterior" -> adata.layers['Psi'], any other mode
        -> the fitted prior: expit(Xc cell_coeff^T + gene_coeff Xg^T + intercept + N(0, sigma)) clipped to +-9 in
        logit space, sigma = adata.varm['sigma'] or `prior_sigma` (simulator.py:17-41).
    effLen : (events, 6) effective lengths, default adata.varm['effLen']; columns 0, 4, 5 are used (:45-51).
    Returns a copy of adata whose count layers are ~ Multinomial(observed total, Phi) (:64-73); like the
    reference it also leaves layers['Psi_sim'] (and 'Psi_sim_noNoise') on the INPUT adata.
    `seed` (not in the reference, whose draws are unseeded) keys both the logit noise and the counter-based
    multinomial draws; default: fresh entropy."""
    import torch
    from .. import _lib
    if seed is None:
        seed = int(np.random.SeedSequence().entropy % (1 << 62))
    if Psi is None and "Psi" not in adata.layers:
        print("Error: no Psi available in adata.layers.")
        sys.exit()
    elif Psi is None:
        if mode == "posterior":
            Psi = np.array(adata.layers['Psi'], np.float32)
        else:
            Z = np.zeros(adata.shape, np.float32)
            if 'Xc' in adata.obsm and adata.obsm['Xc'].shape[1] > 0:
                Z += np.dot(adata.obsm['Xc'], adata.varm['cell_coeff'].T)
            if 'Xg' in adata.varm and adata.varm['Xg'].shape[1] > 0:
                Z += np.dot(adata.obsm['gene_coeff'], adata.varm['Xg'].T)
            if 'intercept' in adata.varm and adata.varm['intercept'].shape[1] > 0:
                Z += adata.varm['intercept'].T
            if 'intercept' in adata.obsm and adata.obsm['intercept'].shape[1] > 0:
                Z += adata.obsm['intercept']
            adata.layers['Psi_sim_noNoise'] = expit(Z)
            sigma = adata.varm['sigma'].T if prior_sigma is None else np.ones([1, adata.shape[1]]) * prior_sigma
            Z += np.random.default_rng(seed).normal(0.0, 1.0, Z.shape) * sigma
            Psi = expit(np.clip(Z, -9, 9)).astype(np.float32)
    adata.layers['Psi_sim'] = Psi

    if effLen is None and 'effLen' not in adata.varm:
        print("Error: no effLen available in adata.varm.")
        sys.exit()
    L = np.asarray(adata.varm['effLen'] if effLen is None else effLen, np.float32)[:, [0, 4, 5]]

    if not torch.cuda.is_available():
        raise RuntimeError("brie_b200: no CUDA device (there is no CPU fallback)")
    lib = _lib.load()
    dev = torch.device(device if device is not None else "cuda")
    Nc, Ng = adata.shape
    ld = _lib.leading_dim(Ng)
    out = adata.copy()
    total = np.zeros((Nc, ld), np.float32)
    for key in layer_keys:
        total[:, :Ng] += _dense(adata.layers[key])
    psi_p = np.zeros((Nc, ld), np.float32)
    psi_p[:, :Ng] = Psi
    eff3 = np.ones((3, ld), np.float32)
    eff3[:, :Ng] = L.T
    with torch.cuda.device(dev):
        d_tot, d_psi, d_eff = (torch.from_numpy(a).to(dev) for a in (total, psi_p, eff3))
        cs = [torch.empty((Nc, ld), dtype=torch.float32, device=dev) for _ in range(3)]
        _lib.check(lib.brie_resample_counts(seed, Nc, Ng, ld, 0, d_tot.data_ptr(), d_psi.data_ptr(), d_eff.data_ptr(),
                                            cs[0].data_ptr(), cs[1].data_ptr(), cs[2].data_ptr(),
                                            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        for key, t in zip(layer_keys[:3], cs):
            out.layers[key] = t[:, :Ng].cpu().numpy()
    return out
