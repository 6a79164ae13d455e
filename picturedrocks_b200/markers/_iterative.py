"""Selector base classes (reference: picturedrocks/markers/_iterative.py:22-83).

`IterativeFeatureSelection` keeps the reference's contract — attributes
`infoset`, `score` (float64 ndarray, -inf for selected features), `S` (list),
methods `add(ind)`, `remove(ind)`, `autoselect(n_feats)`.  The generic
`autoselect` below is the reference's host loop (np.argmax then add,
_iterative.py:64-66); the device-resident selectors override it with a loop
that never leaves the GPU.
"""
from abc import ABC, abstractmethod

import numpy as np


class IterativeFeatureSelection(ABC):
    """Abstract greedy feature selector."""

    @abstractmethod
    def __init__(self, infoset):
        self.infoset = infoset
        self.score = None
        self.S = []

    @abstractmethod
    def add(self, ind):
        """Select feature `ind`."""

    @abstractmethod
    def remove(self, ind):
        """Un-select feature `ind`."""

    def autoselect(self, n_feats):
        """Greedily select `n_feats` features: highest score first, lowest index on ties."""
        for _ in range(n_feats):
            self.add(int(np.argmax(self.score)))


class UnivariateMixin:
    """Selection without redundancy penalty (_iterative.py:69-83).  The score vector is
    a host ndarray; tie order of the top-n is numpy's, obtained with the same two numpy
    calls the reference makes."""

    def add(self, ind):
        self.S.append(ind)
        self.score[ind] = -np.inf

    def remove(self, ind):
        self.S.remove(ind)
        self.score[ind] = self.base_score[ind]

    def autoselect(self, n_feats):
        top = np.argpartition(self.score, -n_feats)[-n_feats:]
        top = top[np.argsort(self.score[top])[::-1]]
        self.S.extend(top)
        self.score[top] = -np.inf
        return top
