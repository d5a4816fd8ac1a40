"""CPU tests of the oracle itself: noise spec known answers, analytic gradients vs
autograd, golden vectors produced by the reference's own base_model.py, closed-form
checks, schedule semantics."""
import os

import numpy as np
import pytest
import torch
from scipy.special import gammaln, logit

from oracle import philox_np as px
from oracle.brie2_oracle import (OracleBRIE2, OracleInit, _Adam, add_pseudo_count, fdr_bh, oracle_fit_matrix,
                                 sigmoid, LEARNING_RATES)
from oracle.brie2_torch_eager import EagerBRIE2
from util import make_problem

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"), allow_pickle=True)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for c, k, want in kat:
        r = px.philox4x32_10(*[np.array([x]) for x in c], k[0], k[1])
        assert " ".join("%08x" % int(x[0]) for x in r) == want


def test_normal_field_moments_and_sharding():
    e = px.normal_field(500, 400, 3, px.PHASE_TRAIN, 1, 42, 3)
    assert abs(e.mean()) < 5e-3 and abs(e.std() - 1) < 5e-3 and abs(np.mean(e ** 4) - 3) < 0.05
    # an event shard sees the same noise as the un-sharded field
    part = px.normal_field(500, 100, 3, px.PHASE_TRAIN, 1, 42, 3, col_offset=250)
    assert np.array_equal(part, e[:, :, 250:350])
    # different step / model / phase -> different streams
    assert not np.array_equal(e, px.normal_field(500, 400, 4, px.PHASE_TRAIN, 1, 42, 3))
    assert not np.array_equal(e, px.normal_field(500, 400, 3, px.PHASE_TRAIN, 2, 42, 3))
    assert not np.array_equal(e, px.normal_field(500, 400, 3, px.PHASE_EVAL, 1, 42, 3))


CASES = [('gene', True, 3, 2, 0, None, None), ('gene', False, 2, 1, 3, None, None),
         ('cell', True, 3, 1, 2, None, None), ('None', True, 3, 0, 0, 0, None),
         ('gene', True, 2, 1, 0, None, 1.5), ('cell', False, 2, 0, 0, None, None)]


@pytest.mark.parametrize("mode,eff,n_layers,Kc,Kg,intercept,sigma", CASES)
def test_marginlik_gradients_match_autograd(mode, eff, n_layers, Kc, Kg, intercept, sigma):
    """target='marginLik' (model_TFProb.py:156-157, 188-189, 202-205): prior-sampled
    log-mean-exp objective; analytic gradients of the prior's parameters vs autograd."""
    Nc, Ng, S = 30, 22, 4
    data, effLen, Xc, Xg = make_problem(Nc, Ng, Kc, Kg, eff, n_layers)
    add_pseudo_count(data, np.float32(0.01))
    om = OracleBRIE2(Nc, Ng, Kc, Kg, effLen, intercept, mode, sigma, dtype=np.float64, seed=7)
    om.Xc, om.Xg = Xc.astype(np.float64), Xg.astype(np.float64)
    eps = om.eps(px.PHASE_TRAIN, 3, S)
    loss, lg, grads = om.loss_and_grads(data, eps, target="marginLik")
    ishape = (Nc, 1) if mode == 'cell' else (1, Ng)
    init = OracleInit(Nc, Ng, Kc, Kg, ishape, ishape, intercept, sigma, seed=7)
    em = EagerBRIE2(Nc, Ng, Kc, Kg, effLen, intercept, mode, sigma, init, torch.float64)
    em.set_design(Xc, Xg)
    cl = [torch.tensor(x, dtype=torch.float64) for x in data]
    te = torch.tensor(eps, dtype=torch.float64)
    tl = em.get_loss(cl, te, target="marginLik")
    vs = em.variables("marginLik")
    assert 'Z_loc' not in vs and sorted(vs) == sorted(grads)
    assert abs(loss - float(tl.detach())) < 1e-9 * abs(loss)
    if vs:
        gs = dict(zip(vs.keys(), torch.autograd.grad(tl, list(vs.values()))))
        for k in gs:
            assert np.abs(grads[k] - gs[k].numpy()).max() < 1e-9 * max(1, np.abs(grads[k]).max()), k
    assert np.abs(lg - em.get_loss(cl, te, axis=0, target="marginLik").detach().numpy()).max() < 1e-9 * np.abs(lg).max()
    # an element without reads contributes exactly nothing (log-mean-exp of zeros)
    zero = [np.zeros_like(x) for x in data]
    l0, lg0, g0 = om.loss_and_grads(zero, eps, target="marginLik")
    assert l0 == 0 and not lg0.any() and all(not np.asarray(v).any() for v in g0.values())


@pytest.mark.parametrize("mode,eff,n_layers,Kc,Kg,intercept,sigma", CASES)
def test_analytic_gradients_match_autograd(mode, eff, n_layers, Kc, Kg, intercept, sigma):
    Nc, Ng, S = 30, 22, 3
    data, effLen, Xc, Xg = make_problem(Nc, Ng, Kc, Kg, eff, n_layers)
    add_pseudo_count(data, np.float32(0.01))
    om = OracleBRIE2(Nc, Ng, Kc, Kg, effLen, intercept, mode, sigma, dtype=np.float64, seed=7)
    om.Xc, om.Xg = Xc.astype(np.float64), Xg.astype(np.float64)
    eps = om.eps(px.PHASE_TRAIN, 3, S)
    loss, lg, grads = om.loss_and_grads(data, eps)
    ishape = (Nc, 1) if mode == 'cell' else (1, Ng)
    init = OracleInit(Nc, Ng, Kc, Kg, ishape, ishape, intercept, sigma, seed=7)
    em = EagerBRIE2(Nc, Ng, Kc, Kg, effLen, intercept, mode, sigma, init, torch.float64)
    em.set_design(Xc, Xg)
    cl = [torch.tensor(x, dtype=torch.float64) for x in data]
    te = torch.tensor(eps, dtype=torch.float64)
    tl = em.get_loss(cl, te)
    vs = em.variables()
    gs = dict(zip(vs.keys(), torch.autograd.grad(tl, list(vs.values()))))
    assert abs(loss - float(tl)) < 1e-9 * abs(loss)
    assert sorted(gs) == sorted(grads)
    for k in gs:
        assert np.abs(grads[k] - gs[k].numpy()).max() < 1e-9 * max(1, np.abs(grads[k]).max()), k
    assert np.abs(lg - em.get_loss(cl, te, axis=0).detach().numpy()).max() < 1e-9 * np.abs(lg).max()


def test_likelihood_matches_reference_base_lik():
    """exp(sum_k c_k log phi_k) * multinomial coefficient == brie.models.base_model.BRIE_base_lik
    (golden, generated from /root/reference/brie/models/base_model.py:20-27)."""
    psi, counts, L, want = GOLD['lik_psi'], GOLD['lik_counts'], GOLD['lik_lengths'], GOLD['lik_value']
    n = psi.size
    effLen = np.zeros((n, 6))
    effLen[:, 0], effLen[:, 4], effLen[:, 5] = L[:, 0], L[:, 1], L[:, 2]
    om = OracleBRIE2(1, n, 0, 0, effLen, 0, 'gene', 1.0, dtype=np.float64, seed=0)
    om.p['Z_loc'] = logit(psi)[None, :]
    om.p['Z_std_log'] = np.full((1, n), -50.0)         # z == Z_loc
    data = [counts[:, k][None, :].astype(np.float64) for k in range(3)]
    # per-element log-lik = -(loss_gene - KL); recover it from two evaluations
    eps = np.zeros((1, 1, n))
    _, lg, _ = om.loss_and_grads(data, eps, want_grads=False)
    _, lg0, _ = om.loss_and_grads([np.zeros_like(d) for d in data], eps, want_grads=False)
    ll = -(lg - lg0)
    coef = gammaln(counts.sum(1) + 1) - gammaln(counts + 1).sum(1)
    assert np.allclose(np.exp(ll + coef), want, rtol=1e-9, atol=1e-300)


def test_ci95_matches_reference_get_ci95():
    """Psi95CI (LogitNormal quantiles, model_TFProb.py:100-106) vs base_model.get_CI95 (1.96)."""
    P, Zs, lo, hi = GOLD['ci_psi'], GOLD['ci_zstd'], GOLD['ci_low'], GOLD['ci_high']
    om = OracleBRIE2(1, P.size, dtype=np.float64)
    om.p['Z_loc'] = logit(P)[None, :]
    om.p['Z_std_log'] = np.log(Zs)[None, :]
    # 1.96 vs 1.959964: d(width) <= 2 * 0.25 * Zs * 3.6e-5
    assert np.all(np.abs(om.Psi95CI[0] - (hi - lo)) <= 2e-5 * Zs + 1e-12)
    assert np.allclose(om.Psi[0], P)


def test_zero_count_element_optimum_is_the_prior():
    """With no reads the ELBO optimum is q = prior: gradients vanish at mu = m, s = sigma, KL = 0."""
    Nc, Ng = 6, 5
    om = OracleBRIE2(Nc, Ng, 0, 0, None, None, 'gene', None, dtype=np.float64, seed=1)
    om.p['sigma_log'][:] = np.log(1.7)
    om.p['Z_loc'][:] = om.p['intercept']
    om.p['Z_std_log'][:] = np.log(1.7)
    z = [np.zeros((Nc, Ng)), np.zeros((Nc, Ng))]
    loss, lg, g = om.loss_and_grads(z, om.eps(0, 0, 2))
    assert abs(loss) < 1e-12
    for k in g:
        assert np.abs(g[k]).max() < 1e-12, k


def test_adam_is_keras_form():
    p = {'x': np.array([1.0, -2.0])}
    ad = _Adam(0.01, p, np.float64)
    ad.apply(p, {'x': np.array([0.5, -3.0])})
    # first step: m = 0.1 g, v = 0.001 g^2, alpha = lr*sqrt(1-b2)/(1-b1)
    alpha = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
    g = np.array([0.5, -3.0])
    want = np.array([1.0, -2.0]) - alpha * 0.1 * g / (np.sqrt(0.001 * g * g) + 1e-7)
    assert np.allclose(p['x'], want, rtol=0, atol=1e-15)


def test_fdr_bh_known_values():
    p = np.array([0.01, 0.04, 0.03, 0.005, 0.5, 1.0])
    want = np.array([0.03, 0.06, 0.06, 0.03, 0.6, 1.0])
    assert np.allclose(fdr_bh(p), want)


def test_fit_schedule_and_recovery():
    """6 stages of int(min_iter/6) steps, only the last stage's trace is kept (+ extensions),
    n_iter counts min_iter (model_TFProb.py:234-258); the fit recovers a planted effect."""
    Nc, Ng = 120, 12
    rng = np.random.default_rng(0)
    x = rng.binomial(1, 0.5, Nc).astype(np.float32)
    psi = np.where(x[:, None] > 0, 0.8, 0.2) * np.ones((1, Ng))
    n = rng.poisson(30, (Nc, Ng))
    c1 = rng.binomial(n, psi).astype(np.float32)
    data = [c1, (n - c1).astype(np.float32)]
    res = oracle_fit_matrix(data, Xc=x[:, None], LRT_index=None, intercept_mode='gene', dtype=np.float32,
                            min_iter=300, max_iter=300, MC_size=2, n_eval=10, seed=3)
    assert res.losses.size == 50 and res.n_iter == [300, 300]
    assert np.abs(res.Psi - psi).mean() < 0.1
    assert np.all(res.ELBO_gain[:, 0] > 5) and np.all(res.fdr[:, 0] < 0.05)
    assert res.cell_coeff.shape == (1, Ng) and np.all(res.cell_coeff > 0.5)


def test_oracle_float32_tracks_float64():
    Nc, Ng = 40, 16
    data, effLen, Xc, _ = make_problem(Nc, Ng, 1, 0, True, 3)
    add_pseudo_count(data, np.float32(0.01))
    out = []
    for dt in (np.float32, np.float64):
        om = OracleBRIE2(Nc, Ng, 1, 0, effLen, None, 'gene', None, dtype=dt, seed=5)
        om.fit(data, Xc=Xc, min_iter=120, max_iter=120, MC_size=3, n_eval=5)
        out.append(om)
    assert np.abs(out[0].Psi - out[1].Psi).max() < 1e-3
    assert np.abs(out[0].loss_gene - out[1].loss_gene).max() <= 1e-4 * np.abs(out[1].loss_gene).max()

PUB = np.load(os.path.join(os.path.dirname(__file__), "golden", "published_lrt.npz"), allow_pickle=True)


@pytest.mark.parametrize("table", ["msEAE", "scNT", "dentate"])
def test_published_reference_tables_pin_pval_and_fdr(table):
    """Outputs of the reference itself (brie-tutorials/*/data/*.brie_ident.tsv, fixture made by
    tests/golden/make_published_lrt.py): pval = chi2.sf(2*ELBO_gain, 1) (model_wrap.py:190) and the
    FDR column = BH applied per fitBRIE batch of ceil(batch_size / n_cells) events (:193-196 under
    :241-256).  Values are printed with 4 significant digits ('%.3e')."""
    from scipy.stats import chi2
    gain, pval, fdr = PUB[table + "_gain"], PUB[table + "_pval"], PUB[table + "_fdr"]
    n_gene = int(np.ceil(int(PUB[table + "_batch_size"]) / int(PUB[table + "_n_cells"])))
    # d ln p / d gain ~ -1, the printed gain is good to 5e-4 relative
    p2 = chi2.sf(2 * gain, df=1)
    assert np.all(np.abs(p2 - pval) <= pval * 1e-3 * (1 + np.abs(gain)))
    assert np.all(pval[gain <= 0] == 1.0)
    got = np.zeros_like(fdr)
    for lo in range(0, len(fdr), n_gene):
        for j in range(fdr.shape[1]):
            got[lo:lo + n_gene, j] = fdr_bh(pval[lo:lo + n_gene, j])
    assert np.all(np.abs(got - fdr) <= 2e-3 * fdr)
    # the pin has teeth: one BH over the whole column does not reproduce the table
    whole = np.stack([fdr_bh(pval[:, j]) for j in range(fdr.shape[1])], 1)
    assert np.mean(np.abs(whole - fdr) > 1e-2 * fdr) > 0.3


def test_kl_term_known_answers_by_numerical_integration():
    """The KL term of get_loss (model_TFProb.py:208) pinned by mathematics, independently of TFP's algebra:
    KL(q || p) = integral q(z) [log q(z) - log p(z)] dz evaluated by adaptive quadrature in float64, for
    posteriors / priors spanning the ranges the fit visits (Z_std = exp(N(0,1)) at init, sigma 0.3..3,
    |mu - m| up to 9).  Also against torch.distributions' closed form (a second, independent library)."""
    from scipy.integrate import quad
    from oracle.brie2_oracle import kl_normal_normal
    rng = np.random.default_rng(5)
    cases = [(0.0, 0.0, 0.0, 0.0), (9.0, np.log(20.0), -2.0, np.log(0.3)), (-9.0, -3.0, 1.0, 1.0)]
    cases += [(rng.uniform(-9, 9), rng.normal(0, 1.2), rng.normal(0, 3), rng.normal(0, 0.6)) for _ in range(40)]
    for mu, lam, m, tau in cases:
        s, sig = np.exp(lam), np.exp(tau)

        def integrand(z):
            lq = -0.5 * ((z - mu) / s) ** 2 - lam - 0.5 * np.log(2 * np.pi)
            lp = -0.5 * ((z - m) / sig) ** 2 - tau - 0.5 * np.log(2 * np.pi)
            return np.exp(lq) * (lq - lp)
        num, _ = quad(integrand, mu - 12 * s, mu + 12 * s, epsabs=1e-13, epsrel=1e-13, limit=400)
        got = float(kl_normal_normal(np.float64(mu), np.float64(lam), np.float64(m), np.float64(tau)))
        assert abs(got - num) <= 1e-9 * max(1.0, abs(num)), (mu, lam, m, tau, got, num)
        t = torch.distributions.kl_divergence(
            torch.distributions.Normal(torch.tensor(mu, dtype=torch.float64), torch.tensor(s, dtype=torch.float64)),
            torch.distributions.Normal(torch.tensor(m, dtype=torch.float64), torch.tensor(sig, dtype=torch.float64)))
        assert abs(got - float(t)) <= 1e-10 * max(1.0, abs(got))
    # and the same term inside the oracle's loss: zero counts => loss_gene is the column sum of the KL
    Nc, Ng = 6, 5
    om = OracleBRIE2(Nc, Ng, 0, 0, None, None, 'gene', None, dtype=np.float64, seed=3)
    zero = [np.zeros((Nc, Ng)), np.zeros((Nc, Ng))]
    _, lg, _ = om.loss_and_grads(zero, np.zeros((1, Nc, Ng)), want_grads=False)
    want = np.zeros(Ng)
    for c in range(Nc):
        for g in range(Ng):
            mu, lam, m, tau = om.p['Z_loc'][c, g], om.p['Z_std_log'][c, g], om.p['intercept'][0, g], om.p['sigma_log'][0, g]
            s, sig = np.exp(lam), np.exp(tau)
            want[g] += quad(lambda z: np.exp(-0.5 * ((z - mu) / s) ** 2 - lam - 0.5 * np.log(2 * np.pi)) * (
                (-0.5 * ((z - mu) / s) ** 2 - lam) - (-0.5 * ((z - m) / sig) ** 2 - tau)),
                mu - 12 * s, mu + 12 * s, epsabs=1e-13, epsrel=1e-13, limit=400)[0]
    assert np.abs(lg - want).max() <= 1e-9 * np.abs(want).max()
