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
