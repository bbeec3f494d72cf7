"""Oracle (TEST INFRASTRUCTURE): multi-segment majority vote.

Restates loops/loops.py:169-247 in plain Python over numpy arrays (integer
work; small cases).  Pinned by tests/golden/vote_cases.npz.
"""
from __future__ import annotations

import numpy as np


def _as_np(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def majority_vote_accuracy(predicted, clip_ids, true_labels, posteriors, tie_strategy="min_label") -> float:
    """Fraction of clips whose voted label equals the clip's (first segment's) label.

    Per clip id in ascending order (np.unique, loops.py:196): count the segment
    predictions; a unique maximum wins (:212-214); on a tie ``"min_label"`` takes
    the smallest tied label (:217-218), ``"max_posterior"`` the label of the
    first segment whose posterior is strictly greatest among segments predicting
    a tied label (:220-232), any other string the tied label seen first in
    segment order (:233-234, dict insertion order of Counter).
    """
    predicted, clip_ids = _as_np(predicted), _as_np(clip_ids)
    true_labels, posteriors = _as_np(true_labels), _as_np(posteriors)
    clips = np.unique(clip_ids)
    hits = 0
    for clip in clips:
        where = np.flatnonzero(clip_ids == clip)
        votes = [int(predicted[i]) for i in where]
        tally = {}
        for v in votes:                      # insertion-ordered like Counter
            tally[v] = tally.get(v, 0) + 1
        top = max(tally.values())
        tied = [lab for lab, cnt in tally.items() if cnt == top]
        if len(tied) == 1:
            winner = tied[0]
        elif tie_strategy == "min_label":
            winner = min(tied)
        elif tie_strategy == "max_posterior":
            best, winner = -np.inf, None
            for k, lab in enumerate(votes):
                if lab in tied and posteriors[where[k]] > best:
                    best, winner = posteriors[where[k]], lab
        else:
            winner = tied[0]
        if winner == int(true_labels[where[0]]):
            hits += 1
    return hits / len(clips)
