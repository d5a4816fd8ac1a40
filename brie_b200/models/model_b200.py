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


class _Normal:
    """What callers read off `tfd.Normal(loc, scale)` (model_TFProb.py:113-127): parameters and `sample`."""

    def __init__(self, loc, scale):
        self.loc, self.scale = _arr(loc), _arr(scale)
        self.parameters = {'loc': self.loc, 'scale': self.scale}

    def mean(self):
        return self.loc

    def stddev(self):
        return _arr(np.broadcast_to(self.scale, self.loc.shape))

    def sample(self, sample_shape=(), seed=None):
        shape = (sample_shape,) if np.isscalar(sample_shape) else tuple(sample_shape)
        rng = np.random.default_rng(seed)
        return _arr(self.loc + self.scale * rng.standard_normal(shape + self.loc.shape).astype(np.float32))


class _LogitNormal(_Normal):
    """What callers read off `tfd.LogitNormal(loc, scale)` (model_TFProb.py:96-106): `quantile`."""

    def quantile(self, q):
        from scipy.special import expit, ndtri
        return _arr(expit(self.loc + self.scale * np.float32(ndtri(q))))

    def sample(self, sample_shape=(), seed=None):
        from scipy.special import expit
        return _arr(expit(super().sample(sample_shape, seed)))


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
        self._fitted_layers = None
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

    @property
    def PsiDist(self):
        """Variational logitNormal distribution of Psi (model_TFProb.py:96-98)."""
        return _LogitNormal(self.Z_loc, self.Z_std)

    @property
    def Z(self):
        """Variational posterior for the logit Psi (model_TFProb.py:113-116)."""
        return _Normal(self.Z_loc, self.Z_std)

    @property
    def Z_prior(self):
        """Predicted informative prior for Z (model_TFProb.py:118-127): N(Xc Wc + Wg Xg^T + intercept, sigma),
        the mean evaluated on the device."""
        e = self._need_fit()
        pm = e.element_terms(0, loglik=False, prior_mean=True)[2]
        return _Normal(pm.cpu().numpy(), self.sigma)

    def _count_tiles(self, count_layers):
        """Device tiles of the caller's count layers (the fitted ones are reused when they are what was fitted)."""
        e = self._need_fit()
        if count_layers is None or count_layers is self._fitted_layers:
            return e.counts
        from .. import ingest
        if len(count_layers) != e.n_layers:
            raise ValueError("BRIE2: %d count layers given, the model was fitted with %d" % (len(count_layers), e.n_layers))
        return [ingest.layer_to_device(x, 0, e.Ng, e.device)[0] for x in count_layers]

    def logLik_MC(self, count_layers, target="ELBO", MC_size=1):
        """Marginal log-likelihood of every element on the variational (target "ELBO") or prior ("marginLik")
        distribution with Monte-Carlo sampling (model_TFProb.py:130-191): (Nc, Ng), fresh noise on every call."""
        e = self._need_fit()
        ll = e.element_terms(0, self._count_tiles(count_layers), mc_size=MC_size, margin=(target == "marginLik"))[0]
        return _arr(ll.cpu().numpy())

    def get_loss(self, count_layers, target="ELBO", axis=None, **kwargs):
        """Loss per gene (axis=0), per cell (axis=1) or in total (model_TFProb.py:194-211): each module is summed
        first, then they are combined, as the reference's docstring demands."""
        import torch
        e = self._need_fit()
        margin = target == "marginLik"
        ll, kl, _ = e.element_terms(0, self._count_tiles(count_layers), mc_size=kwargs.get('MC_size', 1),
                                    margin=margin, kl=not margin)
        red = (lambda t: t.sum(dtype=torch.float64)) if axis is None else (lambda t: t.sum(dim=axis, dtype=torch.float64))
        out = -red(ll) if margin else red(kl) - red(ll)
        return _arr(out.to(torch.float32).cpu().numpy())

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
        self._fitted_layers = count_layers
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
