"""Public names of picturedrocks/markers/__init__.py:18-29 that belong to the
mutual-information path (the interactive UI is out of scope)."""
from . import mutualinformation  # noqa: F401
from ._iterative import IterativeFeatureSelection

# BASELINE.json's north_star calls the selector base class "UniversalFeatureSelector"; the
# reference has no type of that name (its base class is IterativeFeatureSelection,
# picturedrocks/markers/_iterative.py:22) — the alias keeps both spellings importable.
UniversalFeatureSelector = IterativeFeatureSelection
from .mutualinformation.infoset import (InformationSet, SparseInformationSet, makeinfoset,
                                        quantile_discretize)
from .mutualinformation.iterative import CIFE, JMI, MIM, CIFEUnsup, UniEntropy
from .mutualinformation import infoset, iterative  # noqa: F401

__all__ = ["makeinfoset", "quantile_discretize", "InformationSet", "SparseInformationSet", "CIFE",
           "CIFEUnsup", "JMI", "MIM", "UniEntropy", "IterativeFeatureSelection",
           "UniversalFeatureSelector"]
