"""Shared helpers for the parity tests (test infrastructure)."""
import ctypes as C

import numpy as np

from oracle import philox_np as px
from oracle.brie2_oracle import OracleBRIE2, OracleInit, add_pseudo_count


def device_eps_provider(seed, model_id, Nc, Ng, col_offset=0):
    """eps(phase, step, S) read back from the device generator, so the oracle sees
    bit-identical noise to the kernels (host libm and the SFU differ by ~3e-6)."""
    import torch
    from brie_b200 import _lib
    lib = _lib.load()

    def eps(phase, step, S):
        out = torch.empty((S, Nc, Ng), dtype=torch.float32, device="cuda")
        _lib.check(lib.brie_philox_normals_device(seed, phase, model_id, step, S, Nc, Ng, col_offset,
                                                  out.data_ptr(),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out.cpu().numpy()
    return eps


def make_problem(Nc, Ng, Kc, Kg, eff=True, n_layers=3, seed=1, design_seed=0):
    from brie_b200.utils.synth import simulate_counts
    d = simulate_counts(Nc, Ng, design='none', seed=seed, with_efflen=eff, n_layers=n_layers)
    rng = np.random.default_rng(design_seed)
    Xc = rng.standard_normal((Nc, Kc)).astype(np.float32)
    if Kc > 0:
        Xc[:, 0] = rng.binomial(1, 0.5, Nc)
    Xg = rng.standard_normal((Ng, Kg)).astype(np.float32)
    data = [x.copy() for x in d['layers']]
    return data, d['effLen'], Xc, Xg


def make_lrt_problem(Nc=150, Ng=48, seed=8):
    """Denser reads and planted covariate effects of mixed strength, so that an LRT at this
    small size produces both significant and non-significant calls."""
    rng = np.random.default_rng(seed)
    x = rng.binomial(1, 0.5, Nc).astype(np.float32)
    beta = rng.choice([0.0, 0.0, 0.8, 2.5], Ng)
    b = rng.normal(0, 1.0, Ng)
    z = b[None, :] + x[:, None] * beta[None, :] + rng.normal(0, 0.5, (Nc, Ng))
    psi = 1 / (1 + np.exp(-z))
    ex = rng.uniform(50, 300, (Ng, 3))
    L1, L2, L3 = ex[:, 1] + 72, np.full(Ng, 72.0), ex[:, 0] + ex[:, 2] - 16
    effLen = np.zeros((Ng, 6), np.float32)
    effLen[:, 0], effLen[:, 2], effLen[:, 4], effLen[:, 5] = L1, L3, L2, L3
    n = rng.poisson(25, (Nc, Ng)) * (rng.uniform(size=(Nc, Ng)) < 0.7)
    p1, p2 = psi * L1, (1 - psi) * L2
    D = p1 + p2 + L3
    c1 = rng.binomial(n, p1 / D)
    c2 = rng.binomial(n - c1, p2 / (D - p1))
    c3 = n - c1 - c2
    return [c1.astype(np.float32), c2.astype(np.float32), c3.astype(np.float32)], effLen, x[:, None], beta


def bar_report(name, diff, bar, envelope=None):
    """Print how many elements of `diff` exceed the north_star bar `bar` (next to the float32-vs-float64 oracle
    envelope when given) and return that count: the tests report violators instead of widening a bar."""
    diff = np.asarray(diff)
    n_viol = int((diff > bar).sum())
    msg = "%s: bar %.0e | max %.2e q99.9 %.2e median %.2e | violators %d of %d (%.4f %%)" % (
        name, bar, diff.max(initial=0), np.quantile(diff, 0.999) if diff.size else 0, np.median(diff) if diff.size else 0,
        n_viol, diff.size, 100.0 * n_viol / max(diff.size, 1))
    if envelope is not None:
        envelope = np.asarray(envelope)
        msg += " | float32-vs-float64 oracle envelope: max %.2e, its own violators %d" % (
            envelope.max(initial=0), int((envelope > bar).sum()))
    print(msg)
    return n_viol
