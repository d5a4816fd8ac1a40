"""Count ingest on the device: sparse / dense host layers -> dense padded
(cells, ld) float32 CUDA tiles, pseudo-count and gene-filter statistics included.

The reference densifies on the host (`.toarray()`, brie/models/model_wrap.py:108-111)
and builds dense float64 matrices for the filter (brie/utils/preprocessing.py:39-45).
Here only the stored counts cross PCIe; everything else is libbrie_b200.so kernels
(csrc/brie_ingest.cu).  torch is used for device memory and the copy stream only.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _round_up(x, m):
    return (x + m - 1) // m * m


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev(x, dtype, device):
    """Host array -> device tensor through pinned memory (asynchronous on the current stream)."""
    a = np.ascontiguousarray(x, dtype=dtype)
    if a.size == 0:
        return torch.empty(0, dtype=torch.from_numpy(a).dtype, device=device)
    t = torch.from_numpy(a)
    try:
        t = t.pin_memory()
    except RuntimeError:            # pinning can fail under memory pressure; pageable copy still works
        pass
    return t.to(device, non_blocking=True)


def layer_to_device(layer, e0, e1, device, ld=None):
    """One count layer, events [e0, e1) -> (Nc, ld) float32 CUDA tensor with zero padding.

    `layer` may be a scipy CSC/CSR (or any scipy sparse) matrix, a numpy array or a torch
    tensor.  Returns (tensor, bytes copied host->device)."""
    from scipy.sparse import issparse, isspmatrix_csc, isspmatrix_csr
    lib = _lib.load()
    device = torch.device(device)
    Nc = int(layer.shape[0])
    n = int(e1 - e0)
    ld = _lib.leading_dim(n) if ld is None else int(ld)
    with torch.cuda.device(device):
        if issparse(layer):
            if not (isspmatrix_csc(layer) or isspmatrix_csr(layer)):
                layer = layer.tocsc()
            out = torch.empty((Nc, ld), dtype=torch.float32, device=device)
            if isspmatrix_csc(layer):
                lo, hi = int(layer.indptr[e0]), int(layer.indptr[e1])
                ptr = _dev(layer.indptr[e0:e1 + 1].astype(np.int64) - lo, np.int64, device)
                idx = _dev(layer.indices[lo:hi], np.int32, device)
                val = _dev(layer.data[lo:hi], np.float32, device)
                _lib.check(lib.brie_ingest_csc(Nc, n, ld, ptr.data_ptr(), idx.data_ptr() if hi > lo else None,
                                               val.data_ptr() if hi > lo else None, out.data_ptr(), _stream(device)))
            else:
                nnz = int(layer.indptr[-1])
                ptr = _dev(layer.indptr, np.int64, device)
                idx = _dev(layer.indices, np.int32, device)
                val = _dev(layer.data, np.float32, device)
                _lib.check(lib.brie_ingest_csr(Nc, int(e0), n, ld, ptr.data_ptr(), idx.data_ptr() if nnz else None,
                                               val.data_ptr() if nnz else None, out.data_ptr(), _stream(device)))
            nbytes = ptr.numel() * 8 + idx.numel() * 4 + val.numel() * 4
            for t in (ptr, idx, val):       # the kernel reads them after this function returns
                t.record_stream(torch.cuda.current_stream(device))
            return out, nbytes
        out = torch.zeros((Nc, ld), dtype=torch.float32, device=device)
        if torch.is_tensor(layer):
            out[:, :n].copy_(layer[:, e0:e1], non_blocking=True)
            return out, (0 if layer.is_cuda else Nc * n * 4)
        src = torch.from_numpy(np.ascontiguousarray(layer[:, e0:e1], dtype=np.float32))
        out[:, :n].copy_(src, non_blocking=True)
        return out, Nc * n * 4


def add_pseudo_count(tiles, pseudo_count):
    """model_wrap.py:113-117 in place on the device tiles (first two layers)."""
    lib = _lib.load()
    c1, c2 = tiles[0], tiles[1]
    with torch.cuda.device(c1.device):
        _lib.check(lib.brie_add_pseudo_count(c1.shape[0], c1.shape[1], float(pseudo_count), c1.data_ptr(),
                                             c2.data_ptr(), _stream(c1.device)))


def gene_stats(tiles, n_events):
    """Per-event filter statistics of device tiles: dict of float64 numpy vectors
    sum1, sum2, sum3, cells_uniq, cells_total (preprocessing.py:38-61)."""
    lib = _lib.load()
    c1, c2 = tiles[0], tiles[1]
    c3 = tiles[2] if len(tiles) > 2 else None
    Nc, ld = c1.shape
    dev = c1.device
    with torch.cuda.device(dev):
        nb = int(lib.brie_gene_stats_scratch_bytes(Nc, ld))
        scratch = torch.empty(nb // 8 + 1, dtype=torch.float64, device=dev)
        stats = torch.empty((5, ld), dtype=torch.float64, device=dev)
        _lib.check(lib.brie_gene_stats(Nc, ld, c1.data_ptr(), c2.data_ptr(), c3.data_ptr() if c3 is not None else None,
                                       stats.data_ptr(), scratch.data_ptr(), _stream(dev)))
    s = stats[:, :n_events].cpu().numpy()
    return dict(sum1=s[0], sum2=s[1], sum3=s[2], cells_uniq=s[3], cells_total=s[4])


def gather_events(tile, keep_idx):
    """Device column subset (`_inplace_subset_var`) of one tile -> new (Nc, ld') tile."""
    lib = _lib.load()
    dev = tile.device
    keep_idx = np.asarray(keep_idx, np.int64)
    n_out = int(keep_idx.size)
    ld_out = max(_lib.leading_dim(n_out), 32)
    with torch.cuda.device(dev):
        src = torch.from_numpy(keep_idx).to(dev)
        out = torch.empty((tile.shape[0], ld_out), dtype=torch.float32, device=dev)
        _lib.check(lib.brie_gather_events(tile.shape[0], tile.shape[1], tile.data_ptr(),
                                          src.data_ptr() if n_out else None, n_out, ld_out, out.data_ptr(),
                                          _stream(dev)))
        src.record_stream(torch.cuda.current_stream(dev))
    return out


def filter_stats_device(layers_uniq, layers_ambg, device, chunk_events=None):
    """Filter statistics of whole (possibly sparse) layers, event chunk by event chunk, so the
    dense tile never exceeds `chunk_events` columns.  Supports the reference's layer lists:
    two unique layers and any number of ambiguous layers (their counts are added up)."""
    Nc, Ng = layers_uniq[0].shape
    device = torch.device(device)
    if chunk_events is None:
        free, _ = torch.cuda.mem_get_info(device)
        chunk_events = max(int(free * 0.5 / (Nc * 4 * 4)), 32)
    chunk_events = max(chunk_events // 32, 1) * 32
    out = {k: np.zeros(Ng) for k in ('sum1', 'sum2', 'sum3', 'cells_uniq', 'cells_total')}
    ambg = None
    for l in layers_ambg:                  # rare: several ambiguous layers are added up first (sparse stays sparse)
        ambg = l if ambg is None else ambg + l
    for e0 in range(0, Ng, chunk_events):
        e1 = min(e0 + chunk_events, Ng)
        tiles = [layer_to_device(l, e0, e1, device)[0] for l in layers_uniq[:2]]
        if ambg is not None:
            tiles.append(layer_to_device(ambg, e0, e1, device)[0])
        st = gene_stats(tiles, e1 - e0)
        for k in out:
            out[k][e0:e1] = st[k]
    return out
