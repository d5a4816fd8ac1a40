"""Chunked writer for the dense (cells, events) outputs of a fit.

The reference concatenates per-batch results with `np.append` (brie/models/model_wrap.py:
55-76) and stores `layers['Psi','Z_std','Psi_95CI']` as dense arrays with the note
"TODO: introduce sparse matrix for this" (:289-292): quadratic copying and, at atlas scale
(1M x 20k: 80 GB per layer), more than host RAM.  Here the destination arrays are
allocated once -- in RAM, or as `.npy` memory maps under `out_dir` -- and every finished
event chunk is written into its column range.  With several ranks and an `out_dir` all
ranks write their own event ranges into the same files (one node, one file system), so no
(cells, events) array ever travels through a collective.

Checkpoint / resume: with an `out_dir` every finished chunk also leaves its small
per-event result (`chunk_<e0>_<e1>.pkl`, written after the big arrays are flushed) next to a
`manifest.json` holding the fit's signature; a later run with `resume=True` and the same
signature reloads those chunks instead of refitting them.  Events are independent and noise /
init are keyed by global event ids, so a resumed fit equals an uninterrupted one bit for bit.
"""
import json
import os
import pickle

import numpy as np

BIG_KEYS = ('Psi', 'Psi95CI', 'Z_std', 'Z_loc')


class LayerStore:
    def __init__(self, n_cells, n_events, out_dir=None, rank=0, world=1, dist=None, keys=BIG_KEYS,
                 signature=None, resume=False):
        self.shape = (int(n_cells), int(n_events))
        self.out_dir, self.rank, self.world, self.dist = out_dir, rank, world, dist
        self.keys = tuple(keys)
        self.ranges = []                    # event ranges this rank has written
        self.arrays = {}
        self.resumed = False
        if out_dir is None:
            for k in self.keys:
                self.arrays[k] = np.empty(self.shape, np.float32)
            return
        os.makedirs(out_dir, exist_ok=True)
        self.resumed = bool(resume) and self._manifest_matches(signature)
        self._barrier()                     # every rank has looked at the old manifest
        if rank == 0 and not self.resumed:
            for f in os.listdir(out_dir):   # stale checkpoints of another fit
                if f.startswith("chunk_") and f.endswith(".pkl"):
                    os.remove(os.path.join(out_dir, f))
            for k in self.keys:
                m = np.lib.format.open_memmap(self.path(k), mode='w+', dtype=np.float32, shape=self.shape)
                del m
            with open(os.path.join(out_dir, "manifest.json"), "w") as fh:
                json.dump(dict(shape=self.shape, keys=self.keys, signature=signature), fh)
        self._barrier()
        for k in self.keys:
            self.arrays[k] = np.load(self.path(k), mmap_mode='r+')

    def _manifest_matches(self, signature):
        try:
            with open(os.path.join(self.out_dir, "manifest.json")) as fh:
                man = json.load(fh)
        except (OSError, ValueError):
            return False
        ok = (signature is not None and man.get("signature") == json.loads(json.dumps(signature))
              and tuple(man.get("shape", ())) == self.shape and tuple(man.get("keys", ())) == self.keys)
        return ok and all(os.path.exists(self.path(k)) for k in self.keys)

    def _chunk_path(self, e0, e1):
        return os.path.join(self.out_dir, "chunk_%d_%d.pkl" % (e0, e1))

    def load_chunk(self, e0, e1):
        """The checkpointed per-event result of events [e0, e1) of a resumed fit, or None."""
        if not self.resumed or not os.path.exists(self._chunk_path(e0, e1)):
            return None
        with open(self._chunk_path(e0, e1), "rb") as fh:
            result = pickle.load(fh)
        self.ranges.append((e0, e1))
        return result

    def path(self, key):
        return os.path.join(self.out_dir, "%s.npy" % key)

    def _barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def put(self, e0, result, checkpoint=True):
        """Move the big arrays of one chunk result into columns [e0, e0 + result.Ng) and leave
        (cells, 0) stubs behind, so BRIE_RV.concate only appends the per-event vectors."""
        e1 = e0 + result.Ng
        for k in self.keys:
            v = getattr(result, k, None)
            if v is not None and v.shape[1] == e1 - e0:      # not yet written through fit_BRIE_matrix's out_sink
                self.arrays[k][:, e0:e1] = v
            setattr(result, k, np.zeros((self.shape[0], 0), np.float32))
        self.ranges.append((e0, e1))
        if self.out_dir is not None and checkpoint:   # big arrays on disk first, then the marker file
            for a in self.arrays.values():
                a.flush()
            tmp = self._chunk_path(e0, e1) + ".tmp"
            with open(tmp, "wb") as fh:
                pickle.dump(result, fh)
            os.replace(tmp, self._chunk_path(e0, e1))

    def finish(self):
        """Make every rank see all columns; returns {key: (cells, events) array}."""
        if self.out_dir is not None:
            for a in self.arrays.values():
                a.flush()
            self._barrier()
            return {k: np.load(self.path(k), mmap_mode='r+') for k in self.keys}
        if self.world > 1:
            mine = [(e0, e1, {k: self.arrays[k][:, e0:e1] for k in self.keys}) for e0, e1 in self.ranges]
            parts = [None] * self.world
            self.dist.all_gather_object(parts, mine)
            for r, p in enumerate(parts):
                if r == self.rank:
                    continue
                for e0, e1, blk in p:
                    for k in self.keys:
                        self.arrays[k][:, e0:e1] = blk[k]
        return self.arrays
