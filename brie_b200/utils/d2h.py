"""Device -> host transfer of the dense (cells, events) output layers.

`layers['Psi', 'Psi_95CI', 'Z_std']` are 12 B per cell x event (model_wrap.py:289-292): 4 GB each at 100k x 10k.
A pageable `tensor.cpu()` moves them at a fraction of the link rate and costs a second host copy into the
destination array.  Here row slabs go through two reusable pinned staging buffers on a side stream: slab i + 1 is
in flight over PCIe / NVLink-C2C while slab i is copied from staging into its column range of the destination
(a RAM array or an `.npy` memory map), so the transfer runs at the pinned rate and overlaps the host copy.
"""
import numpy as np
import torch

SLAB_BYTES = 64 << 20
_staging = {}


def _buffers(device, n_floats):
    key = (device.index,)
    cur = _staging.get(key)
    if cur is None or cur[0].numel() < n_floats:
        cur = [torch.empty(n_floats, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        _staging[key] = cur
    return cur


def to_host_columns(src, dst, col0=0, after=None):
    """dst[:, col0:col0 + n] = src, src a (rows, n) float32 device tensor (any row stride), dst a host array.
    `after`: CUDA event the producer of `src` recorded (default: everything enqueued on the current stream so far)."""
    rows, n = src.shape
    if rows == 0 or n == 0:
        return
    dev = src.device
    per = max(1, min(rows, SLAB_BYTES // (n * 4)))
    bufs = _buffers(dev, per * n)
    side = torch.cuda.Stream(dev)
    if after is not None:
        side.wait_event(after)
    else:
        side.wait_stream(torch.cuda.current_stream(dev))
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    slabs = [(r0, min(r0 + per, rows)) for r0 in range(0, rows, per)]

    def issue(i):
        r0, r1 = slabs[i]
        with torch.cuda.stream(side):
            bufs[i % 2][:(r1 - r0) * n].view(r1 - r0, n).copy_(src[r0:r1], non_blocking=True)
            evs[i % 2].record(side)

    issue(0)
    for i, (r0, r1) in enumerate(slabs):
        if i + 1 < len(slabs):
            issue(i + 1)                       # other staging buffer: its previous slab was consumed last iteration
        evs[i % 2].synchronize()
        dst[r0:r1, col0:col0 + n] = bufs[i % 2][:(r1 - r0) * n].view(r1 - r0, n).numpy()
    src.record_stream(side)


def to_host(src):
    """A fresh (rows, n) float32 numpy array holding `src`."""
    out = np.empty(tuple(src.shape), np.float32)
    to_host_columns(src, out, 0)
    return out


class BackgroundCopy:
    """Device -> host copies of a finished event chunk's layers on a worker thread, so that they overlap the next
    chunk's fit (the staging slabs are shared: one BackgroundCopy at a time, `wait()` before the next starts)."""

    def __init__(self, jobs, device):
        import threading
        self.error = None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))      # the producers of every src are enqueued before this point

        def run():
            try:
                torch.cuda.set_device(device)
                for src, dst, col0 in jobs:
                    to_host_columns(src, dst, col0, after=ev)
            except BaseException as e:                     # surfaced by wait()
                self.error = e
        self.thread = threading.Thread(target=run, name="brie-d2h", daemon=True)
        self.thread.start()

    def wait(self):
        self.thread.join()
        if self.error is not None:
            raise self.error
