"""Minimal AnnData stand-in, used only when the `anndata` package is absent.

Carries exactly what the brie-quant path touches: `.X`, `.obs` / `.var` (pandas),
`.obsm`, `.varm`, `.layers`, `.uns`, `.shape`, `.n_obs`, `.n_vars`, boolean / slice
indexing on both axes, `.copy()`, `_inplace_subset_var` (brie/utils/preprocessing.py:63)
and a `.write_h5ad`-like writer (compressed npz, since h5py is not available either).
"""
import numpy as np
import pandas as pd


def _take(m, idx, axis):
    if m is None:
        return None
    if isinstance(m, pd.DataFrame):
        return m.iloc[idx] if axis == 0 else m
    return m[idx] if axis == 0 else m[:, idx]


class AnnDataLite:
    def __init__(self, X=None, obs=None, var=None, obsm=None, varm=None, layers=None, uns=None):
        if X is None and layers:
            X = next(iter(layers.values()))
        self.X = X
        n_obs, n_vars = X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=[str(i) for i in range(n_obs)])
        self.var = var if var is not None else pd.DataFrame(index=[str(i) for i in range(n_vars)])
        self.obsm = dict(obsm or {})
        self.varm = dict(varm or {})
        self.layers = dict(layers or {})
        self.uns = dict(uns or {})

    @property
    def shape(self):
        return self.X.shape

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    def copy(self):
        cp = lambda d: {k: (v.copy() if hasattr(v, 'copy') else v) for k, v in d.items()}
        return AnnDataLite(self.X.copy(), self.obs.copy(), self.var.copy(), cp(self.obsm), cp(self.varm),
                           cp(self.layers), cp(self.uns))

    def _subset(self, rows, cols):
        X, obs, var = self.X, self.obs, self.var
        obsm, varm, layers = self.obsm, self.varm, self.layers
        if rows is not None:
            X = X[rows]
            obs = obs.iloc[rows]
            obsm = {k: _take(v, rows, 0) for k, v in obsm.items()}
            layers = {k: v[rows] for k, v in layers.items()}
        if cols is not None:
            X = X[:, cols]
            var = var.iloc[cols]
            varm = {k: _take(v, cols, 0) for k, v in varm.items()}
            layers = {k: v[:, cols] for k, v in layers.items()}
        return AnnDataLite(X, obs.copy(), var.copy(), obsm, varm, layers, dict(self.uns))

    @staticmethod
    def _norm(sel, n):
        if isinstance(sel, slice):
            return None if sel == slice(None) else np.arange(n)[sel]
        sel = np.asarray(sel)
        return np.where(sel)[0] if sel.dtype == bool else sel

    def __getitem__(self, key):
        rows, cols = key if isinstance(key, tuple) else (key, slice(None))
        return self._subset(self._norm(rows, self.n_obs), self._norm(cols, self.n_vars))

    def _inplace_subset_var(self, idx):
        sub = self._subset(None, self._norm(idx, self.n_vars))
        self.__dict__.update(sub.__dict__)

    def __repr__(self):
        return ("AnnDataLite object with n_obs x n_vars = %d x %d\n    obs: %s\n    var: %s\n    uns: %s\n"
                "    obsm: %s\n    varm: %s\n    layers: %s" % (
                    self.n_obs, self.n_vars, list(self.obs.columns), list(self.var.columns), list(self.uns),
                    list(self.obsm), list(self.varm), list(self.layers)))

    # ---- persistence (npz; `anndata`/`h5py` are not installable in this image) ----
    @staticmethod
    def _put(d, key, v):
        """One array-like into the npz dict: sparse matrices stay sparse (CSC / CSR components), memory-mapped
        (cells, events) output layers (fitBRIE out_dir) are referenced by path -- densifying or embedding either
        costs 80 GB per matrix at atlas scale -- everything else as a dense array."""
        from scipy.sparse import issparse
        if issparse(v):
            v = v.tocsc() if v.format not in ('csc', 'csr') else v
            d[key + "@sparse/data"], d[key + "@sparse/indices"], d[key + "@sparse/indptr"] = v.data, v.indices, v.indptr
            d[key + "@sparse/meta"] = np.array([v.shape[0], v.shape[1], 0 if v.format == 'csc' else 1], np.int64)
        elif isinstance(v, np.memmap) and v.filename is not None:
            d[key + "@memmap"] = np.asarray(str(v.filename))
        else:
            d[key] = np.asarray(v)

    @staticmethod
    def _get_all(z, prefix):
        """{name: array} of everything stored under `prefix/` by _put."""
        from scipy.sparse import csc_matrix, csr_matrix
        out = {}
        for k in z.files:
            if not k.startswith(prefix + "/"):
                continue
            name = k[len(prefix) + 1:]
            if name.endswith("@memmap"):
                out[name[:-7]] = np.load(str(z[k]), mmap_mode='r')
            elif name.endswith("@sparse/meta"):
                base = k[:-len("/meta")]
                r, c, fmt = (int(x) for x in z[k])
                cls = csc_matrix if fmt == 0 else csr_matrix
                out[name[:-len("@sparse/meta")]] = cls((z[base + "/data"], z[base + "/indices"], z[base + "/indptr"]), shape=(r, c))
            elif "@sparse/" not in name:
                out[name] = z[k]
        return out

    def write_npz(self, path):
        d = {"obs_index": np.asarray(self.obs.index, dtype=str), "var_index": np.asarray(self.var.index, dtype=str)}
        self._put(d, "main/X", self.X)
        for c in self.obs.columns:
            d["obs/" + c] = np.asarray(self.obs[c])
        for c in self.var.columns:
            d["var/" + c] = np.asarray(self.var[c])
        for grp, dic in (("obsm", self.obsm), ("varm", self.varm), ("layers", self.layers)):
            for k, v in dic.items():
                self._put(d, grp + "/" + k, v)
        d["uns"] = np.array(self.uns, dtype=object)
        np.savez_compressed(path, **d)

    write_h5ad = write_npz

    @classmethod
    def read_npz(cls, path):
        z = np.load(path, allow_pickle=True)
        obs = pd.DataFrame({k[4:]: z[k] for k in z.files if k.startswith("obs/")}, index=z["obs_index"])
        var = pd.DataFrame({k[4:]: z[k] for k in z.files if k.startswith("var/")}, index=z["var_index"])
        X = cls._get_all(z, "main")["X"] if any(k.startswith("main/") for k in z.files) else z["X"]
        return cls(X, obs, var, cls._get_all(z, "obsm"), cls._get_all(z, "varm"), cls._get_all(z, "layers"),
                   z["uns"].item())
