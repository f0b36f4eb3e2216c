"""mIoU evaluation with the reference's semantics (evaluation.py:6-62): per-sample intersection / union counts,
`score()` = I / max(U, 1) per class.  `EvaluatorIoU.sample(truth, prediction)` is the reference's CPU/numpy call;
`EvaluatorIoU.sample_logits(logits, truth)` (new) takes the network output on the GPU and accumulates the same counts
with ONE fused argmax + confusion-matrix kernel (b2_argmax_confusion) and no host synchronisation until `score()`."""
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

    def sample_logits(self, logits, truth, ignore_value=None):
        """Device path: logits (N,C,H,W) fp32 CUDA tensor, truth (N,1,H,W) / (N,H,W) int64 CUDA tensor.  Equivalent to
        `sample(truth[i], argmax(logits[i]))` for every i; counts stay on the device until `score()` / `flush()`."""
        import torch
        from cutmix_semisup_seg_b200 import ops
        if self.fill_holes:
            raise ValueError('fill_holes needs the CPU path: use sample() on the argmax map')
        if logits.shape[1] != self.num_classes:
            raise ValueError('logits have {} classes, evaluator {}'.format(logits.shape[1], self.num_classes))
        if getattr(self, '_cm_dev', None) is None or self._cm_dev.device != logits.device:
            self.flush()
            self._cm_dev = torch.zeros((self.num_classes * self.num_classes,), dtype=torch.int64, device=logits.device)
        ops.default_backend().argmax_confusion(logits, truth.to(torch.int64), self._cm_dev, ignore_value=ignore_value)

    def flush(self):
        """Fold the device-side confusion matrix into the numpy accumulators (one D2H copy of C*C integers)."""
        cm_dev = getattr(self, '_cm_dev', None)
        if cm_dev is None:
            return
        cm = cm_dev.cpu().numpy().reshape(self.num_classes, self.num_classes)
        self._cm_dev = None
        diag = np.diag(cm)
        self.intersection += diag                                   # (pred == c) & (truth == c)
        self.union += cm.sum(axis=1) + cm.sum(axis=0) - diag         # (pred == c) | (truth == c) over valid pixels
        self.cm += cm

    def score(self):
        self.flush()
        return self.intersection.astype(float) / np.maximum(self.union.astype(float), 1.0)
