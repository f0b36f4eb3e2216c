"""mIoU evaluation with the reference's semantics (evaluation.py:6-62): per-sample intersection / union counts
accumulated on the CPU, `score()` = I / max(U, 1) per class."""
import numpy as np


def fast_cm(tru, pred, num_classes):
    """Confusion matrix through a single bincount."""
    return np.bincount(tru * num_classes + pred, minlength=num_classes * num_classes).reshape((num_classes, num_classes))


def per_class_i_and_u_cm(pred, tru, num_classes, ignore_value=None):
    valid = tru != ignore_value
    inter = np.zeros((num_classes,), dtype=np.int64)
    union = np.zeros((num_classes,), dtype=np.int64)
    for c in range(num_classes):
        p = pred == c
        t = tru == c
        if ignore_value is not None:
            p = p & valid
            t = t & valid
        inter[c] = (p & t).sum()
        union[c] = (p | t).sum()
    return inter, union, fast_cm(tru[valid], pred[valid], num_classes)


class EvaluatorIoU(object):
    def __init__(self, num_classes, fill_holes=False):
        if fill_holes and num_classes != 2:
            raise ValueError('num_classes must be 2 if fill_holes is True')
        self.num_classes, self.fill_holes = num_classes, fill_holes
        self.intersection = np.zeros((num_classes,))
        self.union = np.zeros((num_classes,))
        self.cm = np.zeros((num_classes, num_classes))

    def sample(self, truth, prediction, ignore_value=None):
        if self.fill_holes:
            from scipy.ndimage import binary_fill_holes
            prediction = binary_fill_holes(prediction != 0).astype(int)
        i, u, cm = per_class_i_and_u_cm(prediction, truth, self.num_classes, ignore_value=ignore_value)
        self.intersection += i
        self.union += u
        self.cm += cm

    def score(self):
        return self.intersection.astype(float) / np.maximum(self.union.astype(float), 1.0)
