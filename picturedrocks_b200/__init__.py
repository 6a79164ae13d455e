"""B200-native PicturedRocks mutual-information marker selection.

Drop-in for the reference's `picturedrocks.markers` hot path:
`makeinfoset`, `InformationSet`, `SparseInformationSet` and the greedy selectors
`MIM`, `JMI`, `CIFE`, `UniEntropy`, `CIFEUnsup` (`IterativeFeatureSelection`).
All computation runs in hand-written sm_100a CUDA kernels (libprb200.so); there
is no CPU fallback.  `read.process_clusts` (label coding) and `performance.FoldTester`
(k-fold evaluation) are the host-side callers either side of that path.
"""
from . import data, markers, performance, read  # noqa: F401

__all__ = ["markers", "read", "performance", "data"]
