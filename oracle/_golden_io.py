"""Writer of the golden fixtures with a re-mint CHECK mode (test infrastructure).

``python oracle/make_golden.py --check`` (same for the other ``make_*_golden.py`` recipes) re-runs the whole recipe
against the reference and, instead of overwriting ``tests/golden/*.npz``, compares every freshly minted array with the
committed one: integer / bool / string arrays exactly, floating point within ``FLOAT_TOL`` of the array's largest
magnitude (two fp32 evaluations of the reference in different thread configurations differ by their own rounding
noise; gradients that are cancelling sums -- the additive-attention bias -- carry the most of it).
"""
from __future__ import annotations

import os
import sys

import numpy as np

CHECK = "--check" in sys.argv
FLOAT_TOL = 2e-3
_worst = {}


def save(path: str, **rec) -> None:
    if not CHECK:
        np.savez_compressed(path, **rec)
        return
    old = dict(np.load(path))
    name = os.path.basename(path)
    assert set(old) == set(rec), (name, sorted(set(old) ^ set(rec)))
    worst = 0.0
    for k, new in rec.items():
        new, ref = np.asarray(new), old[k]
        assert new.shape == ref.shape, (name, k, new.shape, ref.shape)
        if k == "oracle_vs_reference_maxrel":
            continue  # a diagnostic of the minting run, not a fixture value
        if new.dtype.kind in "fc":
            scale = max(float(np.abs(ref).max()) if ref.size else 0.0, 1e-30)
            err = float(np.abs(new.astype(np.float64) - ref.astype(np.float64)).max()) / scale if ref.size else 0.0
            assert err <= FLOAT_TOL, (name, k, err)
            worst = max(worst, err)
        else:
            assert np.array_equal(new, ref), (name, k)
    _worst[name] = worst
    print(f"[check] {name}: {len(rec)} arrays reproduce the committed fixture (worst float rel {worst:.1e})")
