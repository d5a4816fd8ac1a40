"""GPU tests of the count ingest: CSC/CSR scatter, pseudo-count, gene-filter
statistics and column gather through the C ABI, bit-exact against the host restatements
(scipy `toarray`, oracle `add_pseudo_count`) and the reference's own filter_genes golden
vectors; sparse input to the fit gives the same result as dense input."""
import os

import numpy as np
import pytest

from brie_b200 import _lib
from oracle.brie2_oracle import add_pseudo_count

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"),
               allow_pickle=True)


def _counts(Nc, Ng, density, seed):
    rng = np.random.default_rng(seed)
    m = rng.poisson(3.0, (Nc, Ng)).astype(np.float32) * (rng.uniform(size=(Nc, Ng)) < density)
    m[:, rng.integers(0, Ng, max(Ng // 10, 1))] = 0          # ragged: some empty event columns
    m[rng.integers(0, Nc, max(Nc // 10, 1)), :] = 0          # and empty cells
    return m


@pytest.mark.parametrize("fmt", ["csc", "csr", "coo"])
@pytest.mark.parametrize("idx_dtype,val_dtype", [(np.int32, np.float32), (np.int64, np.float64)])
def test_sparse_scatter_equals_toarray(fmt, idx_dtype, val_dtype):
    import scipy.sparse as sp
    from brie_b200.ingest import layer_to_device
    Nc, Ng = 203, 157
    dense = _counts(Nc, Ng, 0.2, 1)
    m = getattr(sp, fmt + "_matrix")(dense.astype(val_dtype))
    if fmt != "coo":
        m.indices = m.indices.astype(idx_dtype)
        m.indptr = m.indptr.astype(idx_dtype)
    for e0, e1 in [(0, Ng), (0, 1), (33, 129), (Ng - 5, Ng)]:
        t, nbytes = layer_to_device(m, e0, e1, "cuda")
        got = t.cpu().numpy()
        assert got.shape == (Nc, _lib.leading_dim(e1 - e0))
        assert np.array_equal(got[:, :e1 - e0], dense[:, e0:e1])
        assert not got[:, e1 - e0:].any(), "padding must hold zero counts"
        assert nbytes > 0


def test_sparse_scatter_duplicates_and_empty():
    import scipy.sparse as sp
    from brie_b200.ingest import layer_to_device
    Nc, Ng = 40, 50
    rows = np.array([0, 0, 0, 5, 39, 39], np.int32)
    cols = np.array([3, 3, 3, 49, 0, 0], np.int32)
    vals = np.array([1, 2, 4, 7, 1, 1], np.float32)
    for ctor in (sp.csc_matrix, sp.csr_matrix):
        m = ctor((Nc, Ng), dtype=np.float32)              # build with duplicates kept (not canonical)
        coo = sp.coo_matrix((vals, (rows, cols)), shape=(Nc, Ng))
        order = np.argsort(cols if ctor is sp.csc_matrix else rows, kind="stable")
        major = (cols if ctor is sp.csc_matrix else rows)[order]
        minor = (rows if ctor is sp.csc_matrix else cols)[order]
        ptr = np.zeros((Ng if ctor is sp.csc_matrix else Nc) + 1, np.int32)
        np.add.at(ptr, major + 1, 1)
        m = ctor((vals[order], minor, np.cumsum(ptr).astype(np.int32)), shape=(Nc, Ng))
        assert m.nnz == 6                                  # duplicates still stored separately
        got = layer_to_device(m, 0, Ng, "cuda")[0].cpu().numpy()[:, :Ng]
        assert np.array_equal(got, coo.toarray())
        assert got[0, 3] == 7 and got[39, 0] == 2
        empty = ctor((Nc, Ng), dtype=np.float32)
        assert not layer_to_device(empty, 0, Ng, "cuda")[0].cpu().numpy().any()


def test_pseudo_count_equals_reference_rule():
    from brie_b200.ingest import add_pseudo_count as dev_pc, layer_to_device
    Nc, Ng = 300, 260
    layers = [_counts(Nc, Ng, 0.15, s) for s in (1, 2, 3)]
    tiles = [layer_to_device(x, 0, Ng, "cuda")[0] for x in layers]
    dev_pc(tiles, 0.01)
    ref = [x.copy() for x in layers]
    add_pseudo_count(ref, np.float32(0.01))                 # model_wrap.py:113-117 restated
    for t, r in zip(tiles, ref):
        assert np.array_equal(t.cpu().numpy()[:, :Ng], r)
        assert not t.cpu().numpy()[:, Ng:].any()


def test_gene_stats_and_filter_against_reference_golden():
    """Device statistics reproduce the reference's filter_genes subset and var columns."""
    from scipy.sparse import csc_matrix, csr_matrix
    from brie_b200.ingest import filter_stats_device
    from brie_b200.utils.anndata_lite import AnnDataLite
    from brie_b200.utils.preprocessing import filter_genes
    l1, l2, l3 = GOLD['fg_l1'], GOLD['fg_l2'], GOLD['fg_l3']
    st = filter_stats_device([csc_matrix(l1), csc_matrix(l2)], [csc_matrix(l3)], "cuda", chunk_events=64)
    assert np.array_equal(st['sum1'], l1.astype(np.float64).sum(0))
    assert np.array_equal(st['sum3'], l3.astype(np.float64).sum(0))
    assert np.array_equal(st['cells_uniq'], ((l1 + l2) > 0).sum(0))
    assert np.array_equal(st['cells_total'], ((l1 + l2 + l3) > 0).sum(0))
    for wrap in (lambda x: x, csc_matrix, csr_matrix):
        ad = AnnDataLite(X=l1 + l2 + l3, layers={'isoform1': wrap(l1), 'isoform2': wrap(l2), 'ambiguous': wrap(l3)})
        out = filter_genes(ad, min_counts=50, min_counts_uniq=10, min_cells_uniq=30, min_MIF_uniq=0.001, copy=True,
                           device="cuda")
        assert out.shape == (80, int(GOLD['fg_subset'].sum()))
        assert np.array_equal(np.asarray(out.var['n_counts']), GOLD['fg_n_counts'])
        assert np.array_equal(np.asarray(out.var['n_counts_uniq']), GOLD['fg_n_counts_uniq'])


def test_gather_events():
    from brie_b200.ingest import gather_events, layer_to_device
    dense = _counts(77, 300, 0.3, 5)
    t = layer_to_device(dense, 0, 300, "cuda")[0]
    keep = np.flatnonzero(np.random.default_rng(0).uniform(size=300) < 0.4)
    got = gather_events(t, keep).cpu().numpy()
    assert np.array_equal(got[:, :keep.size], dense[:, keep]) and not got[:, keep.size:].any()
    assert gather_events(t, np.zeros(0, np.int64)).shape == (77, 32)


def test_fit_from_sparse_layers_equals_dense():
    """fit_BRIE_matrix on CSC layers (no host densification) == the same fit on dense arrays,
    and the caller's sparse matrices are left untouched."""
    from scipy.sparse import csc_matrix
    from brie_b200.models import fit_BRIE_matrix
    from util import make_lrt_problem
    data, effLen, Xc, _ = make_lrt_problem(90, 40, seed=5)
    kw = dict(Xc=Xc, effLen=effLen, LRT_index=None, intercept_mode='gene', seed=2, min_iter=300, max_iter=800,
              MC_size=2, n_eval=10)
    dense = fit_BRIE_matrix([x.copy() for x in data], **kw)
    sparse_in = [csc_matrix(x) for x in data]
    sp_res = fit_BRIE_matrix(list(sparse_in), **kw)
    for k in ('Psi', 'Psi95CI', 'Z_std', 'loss_gene', 'ELBO_gain', 'pval', 'fdr', 'losses', 'cell_coeff'):
        assert np.array_equal(getattr(dense, k), getattr(sp_res, k)), k
    assert all(np.array_equal(s.toarray(), x) for s, x in zip(sparse_in, data))
    assert sp_res.h2d_bytes == sum(8 * (s.shape[1] + 1) + 8 * s.nnz for s in sparse_in)
