"""Id alignment for the -c/-g feature files (semantics of brie/utils/base_utils.py:5-59)."""
import numpy as np


def match(ref_ids, new_ids, uniq_ref_only=True):
    """For every id in `ref_ids` the index of the same id in `new_ids`, or None.

    `new_ids` is expected to hold unique ids.  With `uniq_ref_only` a `new_ids` entry is
    handed out once: a repeated reference id after its first occurrence maps to None.
    Returns an object array of len(ref_ids) (ints and None), like the reference, so that
    `.astype(float)` turns the misses into NaN (brie/bin/quant.py:53-55).
    """
    where = {}
    for j, v in enumerate(new_ids):
        where.setdefault(v, j)
    out = np.empty(len(ref_ids), dtype=object)
    used = set()
    for i, v in enumerate(ref_ids):
        j = where.get(v)
        if j is not None and uniq_ref_only and j in used:
            j = None
        if j is not None:
            used.add(j)
        out[i] = j
    return out
