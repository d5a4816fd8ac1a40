"""Size-independent properties of the fused fit at BASELINE.json's C2 size (5 000 x 5 000 -- too
large for the CPU oracle) and oracle parity on ragged / degenerate shapes.

Properties used (all follow from the reference's model, model_TFProb.py:118-211):
  * events are independent when the intercept is per event and Kg = 0 (every parameter is per
    event, model_wrap.py:241-260 relies on the same fact) => fitting an event range alone, with the
    global event offset in the noise counters, reproduces that range of the whole fit;
  * LRT refits are independent models => batching them in one launch changes nothing;
  * an element without reads has zero likelihood gradient => its Z_loc / Z_std_log
    gradient is the closed-form KL gradient, whatever the noise;
  * same seed => same result, bit for bit (no atomics anywhere on the path).
"""
import numpy as np
import pytest

from oracle import philox_np as px
from oracle.brie2_oracle import OracleBRIE2, add_pseudo_count

from util import device_eps_provider, make_problem

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _c2_problem():
    from brie_b200.utils.synth import simulate_counts_device
    return simulate_counts_device(5000, 5000, design='binary1', seed=2, with_efflen=True, n_layers=3)


def _run(sim, cols=None, masks=([0], []), model_ids=None, steps=12, seed=5):
    """12 steps (2 with the loss trace) of a fresh engine on all events or on the event range `cols`."""
    from brie_b200.engine import FitEngine
    lo, hi = (0, 5000) if cols is None else cols
    layers = sim['layers'] if cols is None else [t[:, lo:hi] for t in sim['layers']]
    eng = FitEngine(layers, effLen=sim['effLen'][lo:hi], Xc=sim['Xc'], masks=[list(m) for m in masks],
                    model_ids=model_ids, MC_size=3, seed=seed, trace_cap=4, event_offset=lo,
                    n_events_total=5000, n_events=hi - lo, group_size=100)
    eng.init_params()
    eng.begin_stage(0.01)
    eng.run_steps(steps - 2)
    eng.run_steps(2, 0)
    torch.cuda.synchronize()
    n = hi - lo
    return dict(Z_loc=eng.Z_loc[:, :, :n].clone(), Z_std_log=eng.Z_std_log[:, :, :n].clone(),
                Wc=eng.Wc[:, :, :n].clone(), b=eng.intercept[:, :n].clone(), tau=eng.sigma_log[:, :n].clone(),
                trace=eng.loss_trace[:, :2, :n].clone(), rows_per_cta=eng.sizes.rows_per_cta)


def _same(a, b, exact):
    for k in ('Z_loc', 'Z_std_log', 'Wc', 'b', 'tau', 'trace'):
        assert a[k].shape == b[k].shape, k
        if a[k].numel() == 0:      # a model without covariates has no Wc rows
            continue
        if exact:
            assert torch.equal(a[k], b[k]), k
        else:       # a different row chunking only reorders float32 partial sums of the per-event gradients
            scale = max(float(a[k].abs().max()), 1.0)
            assert float((a[k] - b[k]).abs().max()) <= 2e-5 * scale, k


def test_c2_full_size_determinism_sharding_and_batching_invariance():
    sim = _c2_problem()
    whole = _run(sim)
    assert torch.isfinite(whole['Z_loc']).all() and torch.isfinite(whole['trace']).all()
    # 1. determinism, bit for bit
    again = _run(sim)
    _same(whole, again, exact=True)
    # 2. event sharding: a group-aligned and a ragged (not a multiple of 32) cut
    for lo, hi in [(0, 2500), (2500, 5000), (1217, 3001)]:
        part = _run(sim, cols=(lo, hi))
        ref = {k: (v[..., lo:hi] if k != 'rows_per_cta' else v) for k, v in whole.items()}
        _same(ref, part, exact=part['rows_per_cta'] == whole['rows_per_cta'])
    # 3. model batching: the full model and the refit fitted alone (same RNG model words)
    for m, mask in enumerate(([0], [])):
        alone = _run(sim, masks=(mask,), model_ids=[m])
        kc = len(mask)
        ref = {k: (v[m:m + 1] if k != 'rows_per_cta' else v) for k, v in whole.items()}
        ref['Wc'], alone['Wc'] = ref['Wc'][:, :kc], alone['Wc'][:, :kc]
        _same(ref, alone, exact=alone['rows_per_cta'] == whole['rows_per_cta'])
    # a different seed does change the result
    other = _run(sim, seed=6)
    assert not torch.equal(other['Z_loc'], whole['Z_loc'])


def test_c2_full_size_zero_count_elements_follow_the_closed_form_kl_gradient():
    """One step from a known state: for every element without reads the Adam first moments are
    0.1 x the KL gradient  d/dZ_loc = (mu - m) / sigma^2,  d/dZ_std_log = s^2 / sigma^2 - 1."""
    from brie_b200.engine import FitEngine
    sim = _c2_problem()
    eng = FitEngine(sim['layers'], effLen=sim['effLen'], Xc=sim['Xc'], masks=[[0]], MC_size=3, seed=9,
                    trace_cap=2, n_events=5000)
    eng.init_params()
    Ng = 5000
    mu, lam = eng.Z_loc[0, :, :Ng].double(), eng.Z_std_log[0, :, :Ng].double()
    Xc = torch.from_numpy(sim['Xc']).to(mu.device).double()
    prior = Xc @ eng.Wc[0, :1, :Ng].double() + eng.intercept[0, :Ng].double()[None, :]
    tau = eng.sigma_log[0, :Ng].double()[None, :]
    g_mu = (mu - prior) * torch.exp(-2 * tau)
    g_lam = torch.exp(2 * (lam - tau)) - 1
    zero = (sim['layers'][0][:, :Ng] + sim['layers'][1][:, :Ng] + sim['layers'][2][:, :Ng]) == 0
    assert 0.7 < float(zero.float().mean()) < 0.95
    eng.begin_stage(0.001)
    eng.run_steps(1)
    torch.cuda.synchronize()
    m_mu, m_lam = 10 * eng.adam_Z[0, 0, :, :Ng].double(), 10 * eng.adam_Z[2, 0, :, :Ng].double()
    assert float(((m_mu - g_mu)[zero]).abs().max()) <= 1e-5 * float(g_mu[zero].abs().max())
    rel = ((m_lam - g_lam)[zero]).abs() / (1 + g_lam[zero].abs())
    assert float(rel.max()) <= 1e-5
    # elements with reads do get a likelihood term
    assert float(((m_mu - g_mu)[~zero]).abs().max()) > 1e-2


EDGE_SHAPES = [(1, 1), (1, 37), (5, 1), (3, 31), (9, 128), (257, 129), (300, 64)]


@pytest.mark.parametrize("Nc,Ng", EDGE_SHAPES)
@pytest.mark.parametrize("mode,Kc,Kg", [('gene', 1, 0), ('cell', 0, 2)])
def test_ragged_and_degenerate_shapes_match_oracle(Nc, Ng, mode, Kc, Kg):
    """Single cells / single events / tiles that end inside a lane vector, a row chunk or a warp,
    an event without any read and a cell without any read: per-event loss and the per-element
    gradients of the first step against the float64 oracle."""
    from brie_b200.engine import FitEngine
    S, seed = 3, 21
    data, effLen, Xc, Xg = make_problem(Nc, Ng, Kc, Kg, True, 3, seed=4)
    for d in data:
        d[:, Ng // 2] = 0                    # an event nobody covers
        d[Nc // 2, :] = 0                    # a cell without reads
    if Nc > 1 and Ng > 1:
        data[0][0, 0] = 7                    # make sure something is non-zero
    add_pseudo_count(data, np.float32(0.01))
    eng = FitEngine(data, effLen=effLen, Xc=Xc, Xg=Xg, intercept_mode=mode, MC_size=S, seed=seed, trace_cap=4)
    eng.init_params()
    om = OracleBRIE2(Nc, Ng, Kc, Kg, effLen, None, mode, None, dtype=np.float64, seed=seed)
    pr = eng.model_params(0)
    om.p['Z_loc'] = eng.Z_loc[0, :, :Ng].cpu().numpy().astype(np.float64)
    om.p['Z_std_log'] = eng.Z_std_log[0, :, :Ng].cpu().numpy().astype(np.float64)
    om.p['Wc_loc'] = pr['Wc_loc'].astype(np.float64)
    om.p['Wg_loc'] = pr['Wg_loc'].astype(np.float64)
    om.p['intercept'] = pr['intercept'].astype(np.float64)
    om.Xc, om.Xg = Xc.astype(np.float64), Xg.astype(np.float64)
    eps = device_eps_provider(seed, 0, Nc, Ng)(px.PHASE_TRAIN, 0, S)
    loss, loss_gene, grads = om.loss_and_grads(data, eps)
    eng.begin_stage(0.001)
    eng.run_steps(1, 0)
    torch.cuda.synchronize()
    tr = eng.loss_trace[0, 0, :Ng].cpu().numpy()
    assert np.all(np.isfinite(tr))
    assert np.abs(tr - loss_gene).max() <= 1e-4 * max(np.abs(loss_gene).max(), 1.0)
    for slot, name in ((0, 'Z_loc'), (2, 'Z_std_log')):
        dev = 10 * eng.adam_Z[slot, 0, :, :Ng].cpu().numpy()
        assert np.abs(dev - grads[name]).max() <= 2e-4 * max(np.abs(grads[name]).max(), 1.0), name
    lg = eng.eval_loss_gene(3)
    assert torch.isfinite(lg).all()
    Psi, CI, Zs = eng.posterior(0)
    assert Psi.shape == (Nc, Ng) and float(Psi.min()) >= 0 and float(Psi.max()) <= 1 and float(CI.min()) >= 0


@pytest.mark.parametrize("Nc,Ng,Kc,group", [(300, 1000, 1, 100), (520, 333, 9, 7), (64, 77, 0, 5), (40, 2100, 2, 1)])
def test_active_block_compaction_is_bit_identical(Nc, Ng, Kc, group, monkeypatch):
    """Extension rounds (model_TFProb.py:250-258) with most reference batches frozen: walking only the
    8-event blocks that hold an active event (brie_fit_set_active_blocks) gives bit for bit what the
    dense walk gives, frozen events keep their state, and each model follows its own active set."""
    from brie_b200.engine import FitEngine
    data, effLen, Xc, _ = make_problem(Nc, Ng, Kc, 0, True, 3, seed=6)
    add_pseudo_count(data, np.float32(0.01))
    masks = [list(range(Kc)), list(range(1, Kc))] if Kc > 0 else [[], []]
    rng = np.random.default_rng(3)

    def run(compact):
        if compact:
            monkeypatch.delenv("BRIE_NO_COMPACT", raising=False)
        else:
            monkeypatch.setenv("BRIE_NO_COMPACT", "1")
        eng = FitEngine(data, effLen=effLen, Xc=Xc, masks=masks, model_ids=[0, 1], MC_size=3, seed=11,
                        trace_cap=8, group_size=group, event_offset=37 * group, n_events_total=Ng + 50 * group)
        eng.init_params()
        eng.begin_stage(0.01)
        eng.run_steps(4)
        before = eng.Z_loc.clone()
        outs = []
        for frac in (0.3, 0.05):
            act = np.random.default_rng(int(frac * 100)).random((2, eng.n_groups)) < frac
            act[1, 0] = True
            eng.set_active_groups(act)
            eng.run_steps(3, 0)
            tr = eng.group_trace(3)
            outs.append((act, tr))
        eng.set_active_groups(np.ones((2, eng.n_groups), bool))
        eng.run_steps(2, 0)
        torch.cuda.synchronize()
        st = dict(Z_loc=eng.Z_loc.clone(), Z_std_log=eng.Z_std_log.clone(), adam=eng.adam_Z.clone(),
                  Wc=eng.Wc.clone(), b=eng.intercept.clone(), tau=eng.sigma_log.clone(),
                  small=eng.adam_small.clone(), trace=eng.loss_trace[:, :2].clone())
        return st, outs, before, eng

    a, outs_a, before, eng = run(True)
    assert eng._blk_ids is None                       # all groups active again: dense walk restored
    b, outs_b, _, _ = run(False)
    for k in a:
        assert torch.equal(a[k][..., :Ng], b[k][..., :Ng]) if a[k].dim() > 1 else torch.equal(a[k], b[k]), k
    for (act_a, tr_a), (act_b, tr_b) in zip(outs_a, outs_b):
        assert np.array_equal(act_a, act_b)
        for m in range(2):                            # traces of the active groups agree exactly
            assert np.array_equal(tr_a[m][act_a[m]], tr_b[m][act_b[m]])
    # the compacted rounds did move the active events
    assert not torch.equal(a['Z_loc'], before)


@pytest.mark.parametrize("Nc,Ng,Kc,group,split", [(300, 1000, 1, 100, False), (520, 333, 9, 7, False),
                                                  (64, 77, 0, 5, False), (40, 2100, 2, 1, True),
                                                  (700, 640, 4, 3, True)])
def test_gathered_extension_round_is_bit_identical(Nc, Ng, Kc, group, split, monkeypatch):
    """Extension rounds on physically gathered columns (FitEngine.run_steps_gathered: per-model dense
    sub-fits of the still-active events, brie_fit_buffers.event_ids) against stepping the whole shard
    in place with the frozen events masked: state, Adam moments, small parameters and the loss trace
    of the active groups agree bit for bit; frozen events are untouched; the step counters (Adam t,
    RNG step word) advance the same; cutting the round into several sub-fits (memory-bound case)
    changes nothing."""
    from brie_b200.engine import FitEngine
    data, effLen, Xc, _ = make_problem(Nc, Ng, Kc, 0, True, 3, seed=8)
    add_pseudo_count(data, np.float32(0.01))
    masks = [list(range(Kc)), list(range(1, Kc))] if Kc > 0 else [[], []]

    def run(gathered):
        monkeypatch.setenv("BRIE_NO_COMPACT", "1")
        eng = FitEngine(data, effLen=effLen, Xc=Xc, masks=masks, model_ids=[0, 1], MC_size=3, seed=12,
                        trace_cap=8, group_size=group, event_offset=11 * group, n_events_total=Ng + 30 * group)
        if split:      # pretend only ~1/3 of the widest active set fits the free memory
            per_col = 2 * Nc * 9 * 4 * 1.05 + 2 * (3 + 64) * 4          # run_steps_gathered's own estimate
            monkeypatch.setattr(eng, "_free_bytes", lambda: int(per_col * 300 / 0.8))    # room for 288 columns
        eng.init_params()
        eng.begin_stage(0.01)
        eng.run_steps(4)
        before = eng.Z_loc.clone()
        outs = []
        for frac in (0.3, 0.05):
            act = np.random.default_rng(int(frac * 100) + 1).random((2, eng.n_groups)) < frac
            act[1, 0] = True
            if frac < 0.1:
                act[0, :] = False                       # a model with nothing left to do
            if gathered:
                assert eng.run_steps_gathered(act, 3, force=True)
            else:
                eng.set_active_groups(act)
                eng.run_steps(3, 0)
            outs.append((act, eng.group_trace(3)))
        eng.set_active_groups(np.ones((2, eng.n_groups), bool))
        eng.run_steps(2, 0)
        torch.cuda.synchronize()
        st = dict(Z_loc=eng.Z_loc.clone(), Z_std_log=eng.Z_std_log.clone(), adam=eng.adam_Z.clone(),
                  Wc=eng.Wc.clone(), b=eng.intercept.clone(), tau=eng.sigma_log.clone(),
                  small=eng.adam_small.clone(), trace=eng.loss_trace[:, :2].clone())
        return st, outs, before, eng

    a, outs_a, before, eng_a = run(True)
    b, outs_b, _, eng_b = run(False)
    assert eng_a.gather_rounds == 2 and eng_b.gather_rounds == 0
    assert eng_a.step_counters() == eng_b.step_counters() == (12, 12)
    for k in a:
        assert torch.equal(a[k][..., :Ng], b[k][..., :Ng]) if a[k].dim() > 1 else torch.equal(a[k], b[k]), k
    for (act_a, tr_a), (act_b, tr_b) in zip(outs_a, outs_b):
        assert np.array_equal(act_a, act_b)
        for m in range(2):
            assert np.array_equal(tr_a[m][act_a[m]], tr_b[m][act_b[m]])
    assert not torch.equal(a['Z_loc'], before)
    assert eng_a.launch_count > eng_b.launch_count    # gathers / scatters are counted


def test_gathered_round_declines_when_it_cannot_help():
    """Shared per-cell parameters, everything still active, or no memory: run_steps_gathered does
    nothing and says so (the caller steps in place)."""
    from brie_b200.engine import FitEngine
    data, effLen, Xc, Xg = make_problem(50, 64, 1, 2, True, 3, seed=2)
    add_pseudo_count(data, np.float32(0.01))
    eng = FitEngine(data, effLen=effLen, Xc=Xc, Xg=Xg, MC_size=2, seed=1, trace_cap=4, group_size=8)
    eng.init_params(); eng.begin_stage(0.01)
    assert not eng.run_steps_gathered(np.zeros((1, eng.n_groups), bool), 2, force=True)   # Kg > 0: shared Wg
    eng = FitEngine(data, effLen=effLen, Xc=Xc, MC_size=2, seed=1, trace_cap=4, group_size=8)
    eng.init_params(); eng.begin_stage(0.01)
    assert not eng.run_steps_gathered(np.ones((1, eng.n_groups), bool), 2, force=True)    # all active
    assert not eng.run_steps_gathered(np.zeros((1, eng.n_groups), bool), 2, force=True)   # nothing active
    act = np.zeros((1, eng.n_groups), bool); act[0, 1] = True
    assert not eng.run_steps_gathered(act, 2)            # 8-event batches = whole blocks: in place is as cheap
    eng._free_bytes = lambda: 1000
    assert not eng.run_steps_gathered(act, 2, force=True)                                 # no memory
    assert eng.step_counters() == (0, 0)


def test_gathered_round_survives_out_of_memory(monkeypatch):
    """The sub-fit size is an estimate; if an allocation fails the round must not die half done: before
    anything ran it is handed back to the in-place walk, later the failing part is halved and retried --
    with the same result bit for bit."""
    from brie_b200 import engine as E
    Nc, Ng, group = 200, 960, 3
    data, effLen, Xc, _ = make_problem(Nc, Ng, 1, 0, True, 3, seed=9)
    add_pseudo_count(data, np.float32(0.01))

    def fresh():
        eng = E.FitEngine(data, effLen=effLen, Xc=Xc, masks=[[0], []], MC_size=3, seed=4, trace_cap=4, group_size=group)
        eng.init_params(); eng.begin_stage(0.01); eng.run_steps(3)
        return eng

    act = np.random.default_rng(0).random((2, Ng // group)) < 0.45      # 400+ active events per model
    ref = fresh()
    assert ref.run_steps_gathered(act, 2, force=True)

    real_init, calls = E._GatheredFit.__init__, []

    def flaky(self, parent, cols, trace_cap):
        calls.append(max(len(c) for c in cols))
        if len(calls) in fail_on:
            raise torch.cuda.OutOfMemoryError("injected")
        real_init(self, parent, cols, trace_cap)

    monkeypatch.setattr(E._GatheredFit, "__init__", flaky)
    fail_on = {1}
    eng = fresh()
    before = eng.Z_loc.clone()
    assert not eng.run_steps_gathered(act, 2, force=True) and eng.step_counters() == (3, 3)
    assert torch.equal(eng.Z_loc, before)
    calls.clear()
    fail_on = {2}
    eng = fresh()
    per_col = 2 * Nc * 9 * 4 * 1.05 + 2 * (2 + 64) * 4
    monkeypatch.setattr(eng, "_free_bytes", lambda: int(per_col * 270 / 0.8))      # 256 columns at a time: two parts
    assert eng.run_steps_gathered(act, 2, force=True)
    assert len(calls) >= 4 and calls[2] < calls[1]                                 # the second part was halved
    torch.cuda.synchronize()
    for a, b in ((eng.Z_loc, ref.Z_loc), (eng.adam_Z, ref.adam_Z), (eng.Wc, ref.Wc), (eng.loss_trace, ref.loss_trace)):
        assert torch.equal(a, b)
    assert eng.step_counters() == ref.step_counters() == (5, 5)
