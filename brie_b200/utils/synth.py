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
