"""Segmentation metrics of the evaluation step (models/lightning_model.py:16-55 of the reference, which
delegates to segeval 2.0.11): get_boundaries, Pk, WindowDiff -- vectorised integer arithmetic on the host
tag vectors, exact rational results returned as Decimal like segeval does.

These are per-episode O(len) integer counts on data that is already on the host (the tag lists returned by
`forward`), so they stay host-side; SURVEY.md section 8f row 2 lists an on-device version as a later step.
"""
from __future__ import annotations

from decimal import Decimal

import numpy as np


def get_boundaries(boundaries):
    """0/1 boundary vector -> list of segment masses (lightning_model.py:16-24)."""
    b = np.asarray(boundaries).astype(bool)
    ends = np.flatnonzero(b)
    if ends.size == 0:
        return []
    return np.diff(np.concatenate(([-1], ends))).tolist()


def _segment_ids(boundaries):
    """Position labels after forcing the last unit to close a segment (compute_Pk sets boundaries[-1] = 1)."""
    b = np.asarray(boundaries).astype(np.int64).copy()
    b[-1] = 1
    return np.concatenate(([0], np.cumsum(b)[:-1])), int(b.sum())


def _window(n, n_segments):
    """segeval default: round-half-even(mean reference mass / 2), at least 2 -- in exact integer arithmetic."""
    den = 2 * n_segments
    q, r = divmod(n, den)
    if 2 * r > den or (2 * r == den and q % 2 == 1):
        q += 1
    return q if q > 1 else 2


def compute_Pk(boundaries, ground_truth, window_size=None):
    hyp, _ = _segment_ids(boundaries)
    ref, nseg = _segment_ids(ground_truth)
    n = len(ref)
    if len(hyp) != n:
        raise ValueError("hypothesis and reference differ in length")
    k = window_size if window_size is not None else _window(n, nseg)
    if n - k <= 0:
        return Decimal(0)
    same_ref = ref[:-k] == ref[k:]
    same_hyp = hyp[:-k] == hyp[k:]
    return Decimal(int(np.count_nonzero(same_ref != same_hyp))) / Decimal(n - k)


def compute_window_diff(boundaries, ground_truth, window_size=None):
    hyp, _ = _segment_ids(boundaries)
    ref, nseg = _segment_ids(ground_truth)
    n = len(ref)
    if len(hyp) != n:
        raise ValueError("hypothesis and reference differ in length")
    k = window_size if window_size is not None else _window(n, nseg)
    if n - k <= 0:
        # segeval asserts here; the reference catches it and substitutes Pk (lightning_model.py:634-637)
        raise AssertionError("window larger than the segmentation")
    # number of boundaries inside a window = difference of segment ids at its two ends
    return Decimal(int(np.count_nonzero((ref[k:] - ref[:-k]) != (hyp[k:] - hyp[:-k])))) / Decimal(n - k)


def f1_boundary(target, pred):
    """sklearn.metrics.f1_score(target, pred, labels=[1], average=None)[0] (0.0 when undefined)."""
    t = np.asarray(target).astype(bool)
    p = np.asarray(pred).astype(bool)
    tp = int(np.count_nonzero(t & p))
    denom = 2 * tp + int(np.count_nonzero(~t & p)) + int(np.count_nonzero(t & ~p))
    return 2.0 * tp / denom if denom else 0.0
