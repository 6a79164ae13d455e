"""Host-side selector base classes (picturedrocks_b200/markers/_iterative.py): no GPU needed."""
import numpy as np
import pytest

from picturedrocks_b200.markers._iterative import IterativeFeatureSelection, UnivariateMixin


class _Fixed(UnivariateMixin, IterativeFeatureSelection):
    def __init__(self, scores):
        self.infoset = None
        self.S = []
        self.base_score = np.asarray(scores, dtype=np.float64)
        self.score = self.base_score.copy()


class _Greedy(IterativeFeatureSelection):
    def __init__(self, scores):
        self.infoset = None
        self.S = []
        self.score = np.asarray(scores, dtype=np.float64).copy()

    def add(self, ind):
        self.S.append(ind)
        self.score[ind] = -np.inf

    def remove(self, ind):
        self.S.remove(ind)


def test_univariate_mixin_follows_the_reference_calls():
    s = [0.5, 2.0, 2.0, -1.0, 3.0, 0.5]
    fs = _Fixed(s)
    top = fs.autoselect(3)
    # the reference's two numpy calls (_iterative.py:79-83)
    ref = np.argpartition(np.array(s), -3)[-3:]
    ref = ref[np.argsort(np.array(s)[ref])[::-1]]
    assert np.array_equal(top, ref) and [int(i) for i in fs.S] == [int(i) for i in ref]
    assert np.all(fs.score[ref] == -np.inf) and fs.score[0] == 0.5
    fs.remove(int(ref[0]))
    assert fs.score[ref[0]] == s[ref[0]] and int(ref[0]) not in fs.S
    fs.add(3)
    assert fs.S[-1] == 3 and fs.score[3] == -np.inf
    with pytest.raises(ValueError):
        fs.remove(0)


def test_generic_autoselect_is_argmax_then_add():
    fs = _Greedy([1.0, 7.0, 7.0, 3.0])
    assert fs.autoselect(3) is None
    assert fs.S == [1, 2, 3]  # first index wins ties
    with pytest.raises(TypeError):
        IterativeFeatureSelection(None)  # abstract


def test_universal_feature_selector_alias():
    import picturedrocks_b200.markers as m

    assert m.UniversalFeatureSelector is m.IterativeFeatureSelection
    assert issubclass(m.CIFE, m.UniversalFeatureSelector)
