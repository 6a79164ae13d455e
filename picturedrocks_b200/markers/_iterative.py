"""Selector base classes (reference: picturedrocks/markers/_iterative.py:22-83).

Contract shared by every selector of this package, identical to the reference's:

* attributes `infoset`, `score` (float64 ndarray over the features; `-inf` marks a selected
  feature) and `S` (Python list of selected feature indices, in selection order);
* `add(ind)` / `remove(ind)` select / un-select one feature;
* `autoselect(n_feats)` selects n_feats more features greedily: highest score first, the
  lowest index among equal scores (np.argmax, _iterative.py:64-66).

The greedy device-resident selectors (mutualinformation/iterative.py) override `autoselect`
with a loop that never leaves the GPU; the univariate ones below keep their score vector on
the host, because their whole selection is two numpy calls.
"""
from abc import ABC, abstractmethod

import numpy as np

NEG_INF = -np.inf


class IterativeFeatureSelection(ABC):
    """Abstract greedy feature selector over an information set."""

    @abstractmethod
    def __init__(self, infoset):
        self.infoset, self.score, self.S = infoset, None, []

    @abstractmethod
    def add(self, ind):
        """Select feature `ind` (an int index into the information set's features)."""

    @abstractmethod
    def remove(self, ind):
        """Un-select feature `ind`; raises ValueError if it is not selected."""

    def autoselect(self, n_feats):
        """Host loop of the reference: n_feats times, take the arg-max of the current score
        vector (first index on ties) and `add` it.  Returns None."""
        remaining = int(n_feats)
        while remaining > 0:
            self.add(int(np.argmax(self.score)))
            remaining -= 1


class UnivariateMixin:
    """Selection by a fixed per-feature score, no redundancy penalty (_iterative.py:69-83):
    MIM and UniEntropy.  `score` / `base_score` are host ndarrays.  The order of the top-n —
    including the order among exactly equal scores — is whatever numpy's `argpartition`
    followed by `argsort` give, so the same two calls are made here."""

    def _mark(self, which, value):
        self.score[which] = value

    def add(self, ind):
        self.S.append(ind)
        self._mark(ind, NEG_INF)

    def remove(self, ind):
        self.S.remove(ind)
        self._mark(ind, self.base_score[ind])

    def autoselect(self, n_feats):
        """Select the n_feats best-scoring features at once; returns their indices, best
        first (unlike the greedy selectors, which return None)."""
        cut = np.argpartition(self.score, -n_feats)[-n_feats:]
        ranked = cut[np.argsort(self.score[cut])[::-1]]
        self.S.extend(ranked)
        self._mark(ranked, NEG_INF)
        return ranked
