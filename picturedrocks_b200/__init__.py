"""B200-native PicturedRocks mutual-information marker selection.

Drop-in for the reference's `picturedrocks.markers` hot path:
`makeinfoset`, `InformationSet`, `SparseInformationSet` and the greedy selectors
`MIM`, `JMI`, `CIFE`, `UniEntropy`, `CIFEUnsup` (`IterativeFeatureSelection`).
All computation runs in hand-written sm_100a CUDA kernels (libprb200.so); there
is no CPU fallback.
"""
from . import markers  # noqa: F401

__all__ = ["markers"]
