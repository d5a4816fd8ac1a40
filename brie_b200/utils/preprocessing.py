"""Gene filter in front of the fit (semantics of brie/utils/preprocessing.py:5-83).

Unlike the reference, which allocates two dense float64 (cells, genes) matrices
(`np.zeros(adata.shape)`, preprocessing.py:39-45), only per-gene column statistics are
accumulated, layer by layer, so sparse layers stay sparse.
"""
import numpy as np


def _col_stats(layers):
    """Per-gene (sum, number of cells with a positive total) over the given layers."""
    from scipy.sparse import issparse
    total = None
    for m in layers:
        total = m.copy() if total is None else total + m
    if total is None:
        return None, None
    if issparse(total):
        s = np.asarray(total.sum(0), dtype=np.float64).reshape(-1)
        n = np.asarray((total > 0).sum(0)).reshape(-1)
    else:
        s = np.asarray(total, dtype=np.float64).sum(0)
        n = (np.asarray(total) > 0).sum(0)
    return s, n


def filter_genes(data, min_counts=0, min_cells=0, min_counts_uniq=0, min_cells_uniq=0,
                 min_MIF_uniq=0.001, uniq_layers=['isoform1', 'isoform2'],
                 ambg_layers=['ambiguous'], copy=False, device=None):
    """Keep genes with enough total / unique counts, expressing cells and minor-isoform
    frequency; adds `n_counts` and `n_counts_uniq` to `adata.var` (preprocessing.py:64-65).

    `device` (not in the reference): a CUDA device -> the column statistics are computed by
    libbrie_b200.so from the stored counts (brie_b200.ingest.filter_stats_device), event chunk
    by event chunk; None -> host sparse arithmetic.  Both give identical masks."""
    from scipy.sparse import issparse
    adata = data.copy() if copy else data
    uniq = [adata.layers[k] for k in uniq_layers]
    ambg = [adata.layers[k] for k in ambg_layers]

    if device is not None:
        from ..ingest import filter_stats_device
        st = filter_stats_device(uniq, ambg, device)
        s1, s2 = st['sum1'], st['sum2']
        u_sum, u_cells = s1 + s2, st['cells_uniq']
        t_sum, t_cells = u_sum + st['sum3'], st['cells_total']
    else:
        def colsum(m):
            return np.asarray(m.sum(0), dtype=np.float64).reshape(-1) if issparse(m) \
                else np.asarray(m, dtype=np.float64).sum(0)
        u_sum, u_cells = _col_stats(uniq)
        t_sum, t_cells = _col_stats(uniq + ambg)
        s1, s2 = colsum(uniq[0]), colsum(uniq[1])

    keep = np.ones(adata.n_vars, dtype=bool)
    keep &= t_sum >= min_counts
    keep &= t_cells >= min_cells
    keep &= u_sum >= min_counts_uniq
    keep &= u_cells >= min_cells_uniq
    keep &= s1 >= min_MIF_uniq * u_sum
    keep &= s2 >= min_MIF_uniq * u_sum

    adata._inplace_subset_var(keep)
    adata.var['n_counts'] = t_sum[keep]
    adata.var['n_counts_uniq'] = u_sum[keep]

    dropped = int(np.sum(~keep))
    if dropped > 0:
        terms = []
        if min_cells > 0:
            terms.append('%d cells with any count' % (min_cells))
        if min_counts > 0:
            terms.append('%d total counts' % (min_counts))
        if min_cells_uniq > 0:
            terms.append('%d cells with unique counts' % (min_cells_uniq))
        if min_counts_uniq > 0:
            terms.append('%d unique counts' % (min_counts_uniq))
        if min_MIF_uniq > 0:
            terms.append('%.4f minor isoform frequency' % (min_MIF_uniq))
        print('Filtered out %d genes with less than ' % (dropped) + " or ".join(terms))
    return adata if copy else None
