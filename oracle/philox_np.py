"""ORACLE (test infrastructure, not product): counter-based MC noise in numpy.

The reference draws its MC noise from TensorFlow's stateful global RNG with no
seed (`brie/models/model_TFProb.py:159`, `tfd.Normal(...).sample(MC_size)`;
init at `:18-31`), so it has no reproducible noise to compare against.  Parity
therefore injects noise: this module is an *independent* numpy restatement of
the noise spec the CUDA path implements in `brie_b200/csrc/brie_philox.h`
(Philox4x32-10, Salmon et al. SC'11, + Box-Muller).  Only tests/, smoke() and
bench.py's cpu_baseline leg may import it.

Spec (must match brie_philox.h):
  key     = (seed & 0xffffffff, seed >> 32)
  counter = (col, row, step, stream)
            stream = phase << 28 | model << 16 | block   (block = sample // 4)
  x0..x3  = philox4x32_10(counter, key)
  u(x)    = ((x >> 9) + 0.5) * 2**-23                    in (0, 1), exact in f32
  pair(a, b): r = sqrt(-2 ln u(a)), t = 2 pi u(b)  ->  (r cos t, r sin t)
  normals = pair(x0, x1) ++ pair(x2, x3);  sample s uses normals[s % 4].
Phases: 0 = training step noise, 1 = loss_gene evaluation noise, 2 = init.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

PHASE_TRAIN = 0
PHASE_EVAL = 1
PHASE_INIT = 2

# parameter ids used as the `step` word in PHASE_INIT
INIT_Z_LOC = 0
INIT_Z_STD_LOG = 1
INIT_WC = 2
INIT_WG = 3
INIT_INTERCEPT = 4


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32-valued arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & MASK32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0          # 32x32 -> 64, fits uint64
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def _u01(x):
    return ((x >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)


def _pair(a, b):
    r = np.sqrt(np.float32(-2.0) * np.log(_u01(a)))
    t = np.float32(6.283185307179586) * _u01(b)
    return r * np.cos(t), r * np.sin(t)


def stream_word(phase, model, block):
    return (int(phase) << 28) | (int(model) << 16) | int(block)


def normals4(col, row, step, phase, model, block, seed):
    """Four float32 normals per (col,row) counter."""
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    x0, x1, x2, x3 = philox4x32_10(col, row, step, stream_word(phase, model, block), k0, k1)
    n0, n1 = _pair(x0, x1)
    n2, n3 = _pair(x2, x3)
    return n0, n1, n2, n3


def normal_field(n_rows, n_cols, step, phase, model, seed, n_samples=1,
                 col_offset=0, row_offset=0):
    """eps[s, r, c] float32 for rows r (cells) and cols c (events).

    `col_offset` is the global index of local column 0, so an event shard sees
    the same noise as the un-sharded problem.
    """
    cols = (np.arange(n_cols, dtype=np.uint64) + np.uint64(col_offset))[None, :]
    rows = (np.arange(n_rows, dtype=np.uint64) + np.uint64(row_offset))[:, None]
    out = np.empty((n_samples, n_rows, n_cols), dtype=np.float32)
    for blk in range((n_samples + 3) // 4):
        n4 = normals4(cols, rows, step, phase, model, blk, seed)
        for j in range(4):
            s = blk * 4 + j
            if s < n_samples:
                out[s] = n4[j]
    return out


def init_field(n_rows, n_cols, param, model, seed, col_offset=0, row_offset=0):
    """N(0,1) init values for one parameter array (PHASE_INIT, sample 0)."""
    return normal_field(n_rows, n_cols, param, PHASE_INIT, model, seed, 1,
                        col_offset, row_offset)[0]
