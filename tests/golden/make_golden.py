#!/usr/bin/env python
"""Generate golden vectors by importing the REFERENCE's own code (run in the build
container, where /root/reference exists; the fixtures travel, the reference does not).

Importable pieces of the reference for this path (the TensorFlow model itself is not:
tensorflow / tensorflow_probability are absent):
  brie/models/base_model.py : BRIE_base_lik (:20-27), get_CI95 (:29-36), LogitNormal (:8-17)
  brie/utils/base_utils.py  : match (:5-59)
  brie/utils/preprocessing.py : filter_genes (:5-83)  (duck-typed AnnData)
Writes tests/golden/reference_vectors.npz.
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/brie"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class MiniAnn:
    """The attribute subset filter_genes touches (preprocessing.py:37-65)."""

    def __init__(self, layers, n_vars):
        self.layers = layers
        self.var = {}
        self.n_vars = n_vars
        self.shape = next(iter(layers.values())).shape

    def copy(self):
        return MiniAnn({k: v.copy() for k, v in self.layers.items()}, self.n_vars)

    def _inplace_subset_var(self, idx):
        self.layers = {k: v[:, idx] for k, v in self.layers.items()}
        self.n_vars = int(np.sum(idx))
        self.shape = next(iter(self.layers.values())).shape
        self.subset = idx.copy()


def main():
    bm = load(os.path.join(REF, "models", "base_model.py"), "ref_base_model")
    bu = load(os.path.join(REF, "utils", "base_utils.py"), "ref_base_utils")
    pp = load(os.path.join(REF, "utils", "preprocessing.py"), "ref_preprocessing")
    rng = np.random.default_rng(2024)
    out = {}

    # 1. multinomial base likelihood: phi = [psi*L1, (1-psi)*L2, L3] / sum
    n = 200
    psi = rng.uniform(0.001, 0.999, n)
    counts = rng.poisson(rng.uniform(0.2, 8, (n, 3))).astype(np.int64)
    counts[counts.sum(1) == 0, 0] = 1
    lengths = np.stack([rng.uniform(50, 400, n), rng.uniform(50, 100, n), rng.uniform(80, 600, n)], 1)
    lik = np.array([bm.BRIE_base_lik(psi[i], counts[i], lengths[i]) for i in range(n)])
    out.update(lik_psi=psi, lik_counts=counts, lik_lengths=lengths, lik_value=lik)

    # 2. 95% interval helper
    P = rng.uniform(0.01, 0.99, 300)
    Zs = np.exp(rng.normal(0, 1, 300))
    lo, hi = bm.get_CI95(P, Zs)
    out.update(ci_psi=P, ci_zstd=Zs, ci_low=lo, ci_high=hi)

    # 3. LogitNormal pdf
    x = rng.uniform(0.01, 0.99, 100)
    loc, scale = 0.7, 1.8
    out.update(ln_x=x, ln_loc=loc, ln_scale=scale, ln_pdf=bm.LogitNormal(loc, scale)._pdf(x))

    # 4. id matching
    ref_ids = np.array(["c%03d" % i for i in rng.permutation(60)])
    new_ids = np.array(["c%03d" % i for i in rng.permutation(80)[:50]])
    idx = bu.match(ref_ids, new_ids)
    out.update(match_ref=ref_ids, match_new=new_ids,
               match_idx=np.array([-1 if v is None else int(v) for v in idx]))

    # 5. gene filter
    Nc, Ng = 80, 120
    lam = np.exp(rng.normal(-1.0, 1.5, Ng))
    l1 = rng.poisson(lam[None, :] * rng.uniform(0, 1, (1, Ng)), (Nc, Ng)).astype(np.float32)
    l2 = rng.poisson(lam[None, :] * rng.uniform(0, 1, (1, Ng)), (Nc, Ng)).astype(np.float32)
    l3 = rng.poisson(lam[None, :] * 2, (Nc, Ng)).astype(np.float32)
    ad = MiniAnn({'isoform1': l1, 'isoform2': l2, 'ambiguous': l3}, Ng)
    res = pp.filter_genes(ad, min_counts=50, min_counts_uniq=10, min_cells_uniq=30, min_MIF_uniq=0.001,
                          uniq_layers=['isoform1', 'isoform2'], ambg_layers=['ambiguous'], copy=True)
    out.update(fg_l1=l1, fg_l2=l2, fg_l3=l3, fg_subset=res.subset,
               fg_n_counts=np.asarray(res.var['n_counts']), fg_n_counts_uniq=np.asarray(res.var['n_counts_uniq']))
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
