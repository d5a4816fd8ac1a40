"""Synthetic splicing count generator (host / numpy).

Follows the reference's generative recipe: psi = logistic(N(., .)) as in
`simulator/simuPSI.py:129-130` (theta = 3), counts ~ Multinomial(n, phi) with
phi proportional to [psi*L1, (1-psi)*L2, L3] as in `brie/models/simulator.py:54-73`,
effective lengths from the SE-event formula `brie/utils/count.py:84-95`
(rlen 76, edge_hang 10, junc_hang 2 -> L = [l2+72, 72, l1+l3-16]).
Sparsity (cell detection rate per event ~ Beta(1.2, 5)) matches the published
result tables' cdr medians.
"""
import numpy as np


def make_design(Nc, kind, rng):
    """Cell covariates per BASELINE config."""
    if kind in (None, 'none'):
        return np.zeros((Nc, 0), np.float32)
    if kind == 'binary1':                         # C2: one binary group
        return rng.binomial(1, 0.5, (Nc, 1)).astype(np.float32)
    if kind == 'mixed3':                          # C3: two binary + one standard normal
        return np.concatenate([rng.binomial(1, 0.5, (Nc, 2)),
                               rng.standard_normal((Nc, 1))], axis=1).astype(np.float32)
    if kind == 'pseudotime':                      # C5
        return rng.uniform(0, 1, (Nc, 1)).astype(np.float32)
    if kind == 'wide15':                          # dentate-gyrus design: detection rate + 14 cluster indicators
        X = np.zeros((Nc, 15), np.float32)            # (brie-tutorials/dentateGyrus/data/dentategyrus_cdr_cluster.tsv)
        X[:, 0] = rng.uniform(0.02, 0.3, Nc)
        X[np.arange(Nc), 1 + rng.integers(0, 14, Nc)] = 1
        return X
    raise ValueError(kind)


def simulate_counts(Nc, Ng, design='none', seed=0, with_efflen=True, n_layers=3,
                    Xg=None, effect_frac=0.1):
    """Returns dict(layers=[c1,c2[,c3]] float32 (Nc,Ng), effLen (Ng,6) or None,
    Xc (Nc,Kc), truth=dict(...))."""
    rng = np.random.default_rng(seed)
    Xc = make_design(Nc, design, rng)
    Kc = Xc.shape[1]
    b = rng.normal(0, 3.0, Ng).astype(np.float32)
    Wc = rng.standard_normal((Kc, Ng)).astype(np.float32)
    Wc *= (rng.uniform(size=(1, Ng)) < effect_frac)
    sig = np.exp(rng.normal(0, 0.5, Ng)).astype(np.float32)
    z = Xc @ Wc + b[None, :] + rng.standard_normal((Nc, Ng)).astype(np.float32) * sig[None, :]
    if Xg is not None:
        Wg = rng.standard_normal((Nc, Xg.shape[1])).astype(np.float32) * 0.3
        z = z + Wg @ Xg.T
    z = np.clip(z, -9, 9)
    psi = 1.0 / (1.0 + np.exp(-z))
    if with_efflen:
        ex = rng.uniform(50, 300, (Ng, 3))
        L1, L2, L3 = ex[:, 1] + 72, np.full(Ng, 72.0), ex[:, 0] + ex[:, 2] - 16
        effLen = np.zeros((Ng, 6), np.float32)
        effLen[:, 0], effLen[:, 2] = L1, L3
        effLen[:, 4], effLen[:, 5] = L2, L3
    else:
        L1 = L2 = np.ones(Ng)
        L3 = np.zeros(Ng)
        effLen = None
    cdr = rng.beta(1.2, 5.0, Ng)
    lam = np.exp(rng.normal(1.0, 1.0, Ng))
    n = rng.poisson(lam[None, :], (Nc, Ng)) * (rng.uniform(size=(Nc, Ng)) < cdr[None, :])
    p1 = psi * L1[None, :]
    p2 = (1 - psi) * L2[None, :]
    p3 = np.broadcast_to(L3[None, :], psi.shape)
    D = p1 + p2 + p3
    c1 = rng.binomial(n, p1 / D)
    rest = n - c1
    c2 = rng.binomial(rest, np.clip(p2 / np.maximum(D - p1, 1e-30), 0, 1))
    c3 = rest - c2
    layers = [c1.astype(np.float32), c2.astype(np.float32)]
    if n_layers > 2:
        layers.append(c3.astype(np.float32))
    return dict(layers=layers, effLen=effLen, Xc=Xc,
                truth=dict(psi=psi.astype(np.float32), Wc=Wc, b=b, sigma=sig))


def simulate_counts_device(Nc, Ng, design='none', seed=0, with_efflen=True, n_layers=3, effect_frac=0.1,
                           pseudo_count=0.01, event_offset=0, device="cuda", Xc=None):
    """Same recipe as simulate_counts, drawn on the GPU (brie_simulate_counts): the per-event
    parameters come from numpy (small), the (cells, events) counts never touch host RAM.  `Xc` overrides the design drawn from `seed`.
    Returns dict(layers=[device tensors (Nc, ld) with the pseudo-count applied], ld, effLen, Xc, truth)."""
    import ctypes as C
    import torch
    from .. import _lib
    lib = _lib.load()
    rng = np.random.default_rng(seed)
    if Xc is None:
        Xc = make_design(Nc, design, rng)
    else:                                  # a design shared by several event blocks (seed = block)
        Xc = np.asarray(Xc, np.float32)
    Kc = Xc.shape[1]
    ld = _lib.leading_dim(Ng)
    pad = lambda v, fill=0.0: np.concatenate([v, np.full(ld - Ng, fill)]).astype(np.float32)
    b = rng.normal(0, 3.0, Ng)
    Wc = rng.standard_normal((Kc, Ng)) * (rng.uniform(size=(1, Ng)) < effect_frac)
    sig = np.exp(rng.normal(0, 0.5, Ng))
    if with_efflen:
        ex = rng.uniform(50, 300, (Ng, 3))
        L = np.stack([ex[:, 1] + 72, np.full(Ng, 72.0), ex[:, 0] + ex[:, 2] - 16])
        effLen = np.zeros((Ng, 6), np.float32)
        effLen[:, 0], effLen[:, 2], effLen[:, 4], effLen[:, 5] = L[0], L[2], L[1], L[2]
        eff3 = np.ones((3, ld), np.float32)
        eff3[:, :Ng] = L
    else:
        effLen, eff3 = None, None
    cdr = rng.beta(1.2, 5.0, Ng)
    lam = np.exp(rng.normal(1.0, 1.0, Ng))
    dev = torch.device(device)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    d_mean, d_sd, d_lam, d_cdr = t(pad(b)), t(pad(sig, 1.0)), t(pad(lam)), t(pad(cdr))
    wc_pad = np.zeros((max(Kc, 1), ld), np.float32)
    wc_pad[:Kc, :Ng] = Wc
    d_Wc, d_Xc = t(wc_pad), t(Xc if Kc > 0 else np.zeros((Nc, 1), np.float32))
    d_eff = t(eff3) if eff3 is not None else None
    outs = [torch.empty((Nc, ld), dtype=torch.float32, device=dev) for _ in range(n_layers)]
    with torch.cuda.device(dev):
        _lib.check(lib.brie_simulate_counts(
            seed, Nc, Ng, ld, event_offset, d_mean.data_ptr(), d_sd.data_ptr(), d_Xc.data_ptr(), d_Wc.data_ptr(), Kc,
            d_eff.data_ptr() if d_eff is not None else None, d_lam.data_ptr(), d_cdr.data_ptr(),
            C.c_float(pseudo_count), outs[0].data_ptr(), outs[1].data_ptr(),
            outs[2].data_ptr() if n_layers > 2 else None, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return dict(layers=outs, ld=ld, effLen=effLen, Xc=Xc,
                truth=dict(Wc=Wc.astype(np.float32), b=b.astype(np.float32), sigma=sig.astype(np.float32),
                           cdr=cdr, lam=lam))
