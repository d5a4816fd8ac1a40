"""`BRIE2`: drop-in for brie.models.BRIE2 (brie/models/model_TFProb.py:35-273).

Same constructor, `fit` signature and attribute protocol that
`fit_BRIE_matrix` / `BRIE_RV` rely on (brie/models/model_wrap.py:18-40, 138-146,
174-187): `.numpy()`-bearing `sigma, intercept, Wc_loc, Wg_loc, Psi, Z_loc, Z_std,
losses, loss_gene`, ndarray `Psi95CI`.  The TensorFlow-Probability backend is
replaced by `FitEngine` -> libbrie_b200.so (sm_100a kernels).
"""
import time

import numpy as np

from ..engine import FitEngine


class _Arr(np.ndarray):
    """ndarray with the `.numpy()` accessor the reference's callers use on tf tensors."""

    def numpy(self):
        return np.asarray(self)


def _arr(x):
    return np.asarray(x).view(_Arr)


class Model_init():
    """Injected initial values (model_TFProb.py:12-31).  Unlike the reference this
    draws from numpy's seeded generator; pass it as `init_obj` for reproducible starts.
    """

    def __init__(self, Nc, Ng, Kc, Kg, intercept_shape, sigma_shape, intercept=None, sigma=None, seed=None):
        rng = np.random.default_rng(seed)
        f32 = np.float32
        self.intercept = (rng.standard_normal(intercept_shape).astype(f32) if intercept is None
                          else np.ones(intercept_shape, f32) * f32(intercept))
        self.sigma = np.ones(sigma_shape, f32) if sigma is None else np.ones(sigma_shape, f32) * f32(sigma)
        self.Z_loc = rng.standard_normal((Nc, Ng)).astype(f32)
        self.Z_std = np.exp(rng.standard_normal((Nc, Ng))).astype(f32)
        self.Wc_loc = rng.standard_normal((Kc, Ng)).astype(f32)
        self.Wg_loc = rng.standard_normal((Nc, Kg)).astype(f32)


class BRIE2():
    """
    Ng : number of genes
    Nc : number of cells
    Kg : number of gene features
    Kc : number of cell features
    """

    def __init__(self, Nc, Ng, Kc=0, Kg=0, effLen=None, intercept=None, intercept_mode='gene',
                 sigma=None, tau_prior=[3, 27], name=None, init_obj=None, seed=0, model_id=0,
                 device=None):
        self.Nc, self.Ng, self.Kc, self.Kg = Nc, Ng, Kc, Kg
        self.effLen = effLen
        self.intercept_mode = intercept_mode
        self._intercept_const, self._sigma_const = intercept, sigma
        self._init_obj = init_obj
        self._seed, self._model_id, self._device = seed, model_id, device
        self._engine = None
        self.Xc = None
        self.Xg = None
        # tau_prior is accepted and unused, as in the reference (model_TFProb.py:44)

    # ---- attribute protocol read by BRIE_RV (model_wrap.py:18-40) ----
    def _need_fit(self):
        if self._engine is None:
            raise RuntimeError("BRIE2: call fit() first (parameters live on the device)")
        return self._engine

    @property
    def Z_loc(self):
        e = self._need_fit()
        return _arr(e.Z_loc[0, :, :e.Ng].cpu().numpy())

    @property
    def Z_std(self):
        return _arr(self._post()[2])

    @property
    def Psi(self):
        return _arr(self._post()[0])

    @property
    def Psi95CI(self):
        return np.asarray(self._post()[1])

    def _post(self):
        e = self._need_fit()
        if self._post_cache is None:
            self._post_cache = [t.cpu().numpy() for t in e.posterior(0)]
        return self._post_cache

    @property
    def sigma(self):
        return _arr(self._need_fit().model_params(0)['sigma'])

    @property
    def intercept(self):
        return _arr(self._need_fit().model_params(0)['intercept'])

    @property
    def Wc_loc(self):
        return _arr(self._need_fit().model_params(0)['Wc_loc'])

    @property
    def Wg_loc(self):
        return _arr(self._need_fit().model_params(0)['Wg_loc'])

    def fit(self, count_layers, Xc=None, Xg=None, target="ELBO", optimizer=None, learn_rate=0.05,
            min_iter=1000, max_iter=5000, add_iter=500, epsilon_conv=1e-2, verbose=True,
            MC_size=1, n_eval=500, **kwargs):
        """Fit the model's parameters (model_TFProb.py:214-273).  `optimizer` and
        `learn_rate` are ignored exactly as the reference ignores them (:228-237).
        target: "ELBO" or "marginLik" (:156-157, 188-189, 202-205)."""
        self.target = target
        start_time = time.time()
        from scipy.sparse import issparse
        layers = [(x.toarray() if issparse(x) else np.asarray(x)).astype(np.float32) for x in count_layers]
        self.Xc, self.Xg = Xc, Xg
        xc = None if (Xc is None or self.Kc == 0) else np.asarray(Xc, np.float32)
        xg = None if (Xg is None or self.Kg == 0) else np.asarray(Xg, np.float32)
        trace_cap = max(int(min_iter / 6), int(add_iter), 1)
        self._engine = FitEngine(layers, effLen=self.effLen, Xc=xc, Xg=xg, masks=None,
                                 model_ids=[self._model_id], intercept=self._intercept_const,
                                 intercept_mode=self.intercept_mode, sigma=self._sigma_const,
                                 MC_size=MC_size, seed=self._seed, device=self._device,
                                 trace_cap=trace_cap, target=target)
        self._post_cache = None
        e = self._engine
        e.fit(min_iter=min_iter, max_iter=max_iter, add_iter=add_iter, epsilon_conv=epsilon_conv,
              n_eval=n_eval, init_objs=None if self._init_obj is None else [self._init_obj])
        self.losses = _arr(e.losses[0])
        self.loss_gene = _arr(e.loss_gene[0].cpu().numpy())
        self.n_iter = int(e.n_iter[0, 0])
        if verbose:
            print("[BRIE2] model fit with %d steps in %.2f min, loss: %.2f" % (
                self.n_iter, (time.time() - start_time) / 60, float(np.sum(self.loss_gene))))
        return self.losses
