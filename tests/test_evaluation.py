"""CPU: the evaluator keeps the reference's semantics (evaluation.py:6-62) and the identity the device path relies on --
per-class intersection / union are functions of the valid-pixel confusion matrix."""
import numpy as np
import pytest

import evaluation


@pytest.mark.parametrize('c,ignore', [(5, 255), (21, 255), (2, None)])
def test_intersection_union_follow_from_confusion_matrix(c, ignore):
    rng = np.random.RandomState(c)
    tru = rng.randint(0, c, size=(37, 41))
    pred = rng.randint(0, c, size=(37, 41))
    if ignore is not None:
        tru[:5] = ignore
    i, u, cm = evaluation.per_class_i_and_u_cm(pred, tru, c, ignore_value=ignore)
    diag = np.diag(cm)
    assert np.array_equal(i, diag)
    assert np.array_equal(u, cm.sum(axis=1) + cm.sum(axis=0) - diag)
    ev = evaluation.EvaluatorIoU(c)
    ev.sample(tru, pred, ignore_value=ignore)
    assert np.allclose(ev.score(), i / np.maximum(u, 1.0))


def test_fill_holes_only_for_binary():
    with pytest.raises(ValueError):
        evaluation.EvaluatorIoU(3, fill_holes=True)
