"""ORACLE (test infrastructure, not product): op-for-op eager restatement.

Mirrors the reference's TensorFlow op graph one op at a time on torch CPU
tensors -- same (S,Nc,Ng[,3]) intermediates, reverse-mode autodiff instead of
analytic gradients, TF-form Adam -- so that it is (a) an independent check of
the analytic gradients in `brie2_oracle.py` and (b) the CPU baseline timed by
bench.py ("restated reference, PyTorch-CPU eager -- not TensorFlow"; the real
TF backend cannot be installed here).

Reference lines restated: brie/models/model_TFProb.py:118-127 (Z_prior),
:130-191 (logLik_MC), :194-211 (get_loss), :214-273 (fit loop).
"""
import math

import numpy as np
import torch

LEARNING_RATES = [0.001, 0.005, 0.01, 0.02, 0.01, 0.005]


class EagerBRIE2:
    def __init__(self, Nc, Ng, Kc=0, Kg=0, effLen=None, intercept=None,
                 intercept_mode='gene', sigma=None, init_obj=None, dtype=torch.float32):
        self.Nc, self.Ng, self.Kc, self.Kg = Nc, Ng, Kc, Kg
        self.dtype = dtype
        self.intercept_mode = intercept_mode
        self.effLen = None if effLen is None else torch.as_tensor(np.asarray(effLen), dtype=dtype)
        ishape = (Nc, 1) if intercept_mode.upper() == 'CELL' else (1, Ng)

        def var(x, shape=None, grad=True):
            t = torch.tensor(np.asarray(x), dtype=dtype)
            if shape is not None:
                t = t.reshape(shape)
            return t.clone().requires_grad_(grad)

        self.intercept = var(init_obj.intercept, ishape, intercept is None)
        self.sigma_log = var(np.log(np.asarray(init_obj.sigma, np.float64)), ishape, sigma is None)
        self.Z_loc = var(init_obj.Z_loc)
        if hasattr(init_obj, 'Z_std_log'):
            self.Z_std_log = var(init_obj.Z_std_log)
        else:
            self.Z_std_log = var(np.log(np.asarray(init_obj.Z_std, np.float64)))
        self.Wc_loc = var(init_obj.Wc_loc, (Kc, Ng), Kc > 0)
        self.Wg_loc = var(init_obj.Wg_loc, (Nc, Kg), Kg > 0)
        self.Xc = None
        self.Xg = None

    def variables(self, target="ELBO"):
        # marginLik never reads the variational parameters (:156-157), so they are not watched
        out = {'Z_loc': self.Z_loc, 'Z_std_log': self.Z_std_log} if target == "ELBO" else {}
        if self.intercept.requires_grad:
            out['intercept'] = self.intercept
        if self.sigma_log.requires_grad:
            out['sigma_log'] = self.sigma_log
        if self.Kc > 0 and self.Xc is not None:
            out['Wc_loc'] = self.Wc_loc
        if self.Kg > 0 and self.Xg is not None:
            out['Wg_loc'] = self.Wg_loc
        return out

    def z_prior_loc(self):                                       # :118-127
        zz = torch.zeros((self.Nc, self.Ng), dtype=self.dtype)
        if self.Kc > 0 and self.Xc is not None:
            zz = torch.matmul(self.Xc, self.Wc_loc)
        if self.Kg > 0 and self.Xg is not None:
            zz = zz + torch.matmul(self.Wg_loc, self.Xg.T)
        return zz + self.intercept

    def logLik_MC(self, count_layers, eps, target="ELBO"):       # :130-191
        if target == "marginLik":                                # Z_prior.sample (:156-157)
            _Z = self.z_prior_loc().unsqueeze(0) + torch.exp(self.sigma_log).unsqueeze(0) * eps
        else:
            Z_std = torch.exp(self.Z_std_log)
            _Z = self.Z_loc.unsqueeze(0) + Z_std.unsqueeze(0) * eps  # Normal.sample, reparameterised
        ls = torch.nn.functional.logsigmoid
        if self.effLen is None:
            ll = count_layers[0].unsqueeze(0) * ls(_Z) + count_layers[1].unsqueeze(0) * ls(0 - _Z)
        else:
            _Z = _Z.unsqueeze(3)
            Psi_logs = torch.cat((ls(_Z), ls(0 - _Z), torch.zeros_like(_Z)), dim=3)
            phi_log = Psi_logs + torch.log(self.effLen[:, [0, 4, 5]])[None, None]
            phi_log = phi_log - torch.logsumexp(phi_log, dim=3, keepdim=True)
            ll = (count_layers[0].unsqueeze(0) * phi_log[:, :, :, 0] +
                  count_layers[1].unsqueeze(0) * phi_log[:, :, :, 1])
            if len(count_layers) > 2:
                ll = ll + count_layers[2].unsqueeze(0) * phi_log[:, :, :, 2]
        if target == "marginLik":                                # tfp.math.reduce_logmeanexp (:188-189)
            return torch.logsumexp(ll, dim=0) - math.log(ll.shape[0])
        return ll.mean(0)

    def kl(self):                                                # tfd.kl_divergence(Normal, Normal)
        a_loc, a_scale = self.Z_loc, torch.exp(self.Z_std_log)
        b_loc, b_scale = self.z_prior_loc(), torch.exp(self.sigma_log)
        diff_log_scale = torch.log(a_scale) - torch.log(b_scale)
        return (0.5 * (a_loc / b_scale - b_loc / b_scale) ** 2 +
                0.5 * torch.expm1(2. * diff_log_scale) - diff_log_scale)

    def get_loss(self, count_layers, eps, axis=None, target="ELBO"):   # :194-211
        if target == "marginLik":                                # :202-205
            ll = self.logLik_MC(count_layers, eps, target)
            return -ll.sum() if axis is None else -ll.sum(axis)
        kl, ll = self.kl(), self.logLik_MC(count_layers, eps)
        if axis is None:
            return kl.sum() - ll.sum()
        return kl.sum(axis) - ll.sum(axis)

    def adam_step(self, state, lr, grads):
        state['t'] += 1
        t = state['t']
        alpha = lr * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        with torch.no_grad():
            for k, v in self.variables().items():
                g = grads[k]
                m = state['m'].setdefault(k, torch.zeros_like(v))
                vv = state['v'].setdefault(k, torch.zeros_like(v))
                m += (g - m) * (1 - 0.9)
                vv += (g * g - vv) * (1 - 0.999)
                v -= (m * alpha) / (torch.sqrt(vv) + 1e-7)
            self.Z_loc.clamp_(-9, 9)
            if self.intercept.requires_grad:
                self.intercept.clamp_(-9, 9)

    def train_step(self, count_layers, eps, state, lr):
        vs = self.variables()
        loss = self.get_loss(count_layers, eps)
        gs = torch.autograd.grad(loss, list(vs.values()))
        self.adam_step(state, lr, dict(zip(vs.keys(), gs)))
        return float(loss.detach())

    def set_design(self, Xc, Xg):
        self.Xc = None if Xc is None else torch.as_tensor(np.asarray(Xc), dtype=self.dtype)
        self.Xg = None if Xg is None else torch.as_tensor(np.asarray(Xg), dtype=self.dtype)
