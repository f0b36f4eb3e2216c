"""Drop-in for the reference's `toy2d/generate_data.py` (BASELINE config 1, the CPU plumbing configuration):
2-D point-classification data sets for `toy2d_train.py`, without scikit-image / batchup (neither is installed here).

Same public names and arithmetic as the reference (file:lines cited per function): `Dataset2D`, `ClassificationDataset2D`,
`SplitClassificationDataset2D`, `ClassificationDatasetFromImage2D`, `classification_dataset_from_image`,
`spiral_classification_dataset`, and the same numpy RandomState draw order, so a seeded data set equals the reference's
(tests/test_toy2d.py pins `classification_dataset_from_image` on the reference's own `curve_mask_v3.png` against the reference's
committed `curve_mask_v3_35.pkl` when /root/reference is present).  The three scikit-image calls are restated:
  img_as_float(rgb2grey(img))  luma 0.2125 R + 0.7154 G + 0.0721 B of the [0,1] image (skimage/color/colorconv.py); a
                               single-channel image is only scaled
  downscale_local_mean(a, f)   block mean (skimage/transform/_warps.py)
  roberts(img)                 sqrt((d1^2 + d2^2) / 2) of the two 2x2 cross differences (skimage/filters/edges.py; plot only)
New: `classification_dataset_from_pickle` loads a data set written by the reference's `generate_data.py clf` command."""
import pickle

import numpy as np

# ImageNet-free toy problem: everything here is host-side numpy (the reference generates it on the CPU as well)


def blend(a, b, t):
    return a + (b - a) * t


def _grey_float(img):
    """skimage.util.img_as_float(skimage.color.rgb2grey(img)) for uint8 / float input."""
    img = np.asarray(img)
    if img.dtype == np.uint8:
        f = img.astype(np.float64) / 255.0
    elif img.dtype == bool:
        f = img.astype(np.float64)
    else:
        f = img.astype(np.float64)
    if f.ndim == 2:
        return f
    rgb = f[..., :3]
    return rgb @ np.array([0.2125, 0.7154, 0.0721])


def _block_mean(a, factors):
    fy, fx = factors
    h, w = a.shape
    ph, pw = (-h) % fy, (-w) % fx
    if ph or pw:
        a = np.pad(a, ((0, ph), (0, pw)), mode='constant')
    return a.reshape(a.shape[0] // fy, fy, a.shape[1] // fx, fx).mean(axis=(1, 3))


def _roberts(img):
    """Roberts cross edge magnitude with skimage's normalisation (each diagonal difference, then sqrt of the mean square)."""
    img = img.astype(np.float64)
    p = np.pad(img, 1, mode='reflect')
    d1 = p[1:-1, 1:-1] - p[2:, 2:]
    d2 = p[1:-1, 2:] - p[2:, 1:-1]
    return np.sqrt((d1 * d1 + d2 * d2) / 2.0)


class Dataset2D(object):
    """Points in [-1, 1]^2 ("real" space) with an image-space view of `img_size` pixels (reference :21-41)."""

    def __init__(self, X, y, img_size):
        self.img_size = img_size
        self.img_scale = np.array(img_size).astype(float)
        self.X, self.y = X, y
        gx, gy = np.meshgrid(np.arange(img_size[1]), np.arange(img_size[0]))
        self.px_grid = np.stack([gy, gx], axis=2) + 0.5

    def load_supervised(self, path):
        raise NotImplementedError('Abstract for {}'.format(type(self)))

    def img_to_real(self, x):
        return (x / self.img_scale) * 2.0 - 1.0

    def real_to_img(self, x):
        return (x + 1.0) * 0.5 * self.img_scale


class ClassificationDataset2D(Dataset2D):
    """Supervised / unsupervised index split plus the sample-density image of the plots (reference :44-104)."""

    def __init__(self, X, y, img_size, sup_indices, unsup_indices):
        super(ClassificationDataset2D, self).__init__(X, y, img_size)
        self.sup_X, self.sup_y = self.X[sup_indices], self.y[sup_indices]
        self.unsup_X, self.unsup_y = self.X[unsup_indices], self.y[unsup_indices]
        self.sup_X_img = self.real_to_img(self.sup_X)
        self.unsup_X_img = self.real_to_img(self.unsup_X)
        X_img = self.real_to_img(X)
        bins = np.arange(self.img_size[0] * 16) / 16.0
        dens, _, _ = np.histogram2d(X_img[:, 0], X_img[:, 1], bins=(bins, bins))
        dens = _block_mean(dens.astype(float), (16, 16)) * 256.0
        self.dens_img = 1.0 - (0.75 ** dens)
        self.px_grid_vis = self.img_to_real(self.px_grid.reshape((-1, 2)))

    def load_supervised(self, path):
        with open(path, 'rb') as f:
            data = pickle.load(f)
        self.sup_X, self.sup_y = data['clf_sup_X'], data['clf_sup_y']
        self.sup_X_img = self.real_to_img(self.sup_X)

    def _plot_base(self, pred_y1, pred_grad, flat_ndim):
        vis = np.zeros(tuple(self.img_size) + (3,), dtype=float)
        vis += 1.0 - self._dens_for_plot()[:, :, None]
        if pred_y1.ndim == flat_ndim or pred_y1.ndim < 2:
            pred_y1 = pred_y1.reshape(self.img_size)
        vis = blend(vis, np.array([[[0.0, 0.75, 0.0]]]), pred_y1[:, :, None] * 0.3)
        if pred_grad is not None:
            if pred_grad.ndim != 2 or pred_grad.shape != tuple(self.img_size):
                pred_grad = pred_grad.reshape(self.img_size)
            pred_grad = np.sqrt(pred_grad / max(abs(pred_grad).max(), 1e-30))
            vis = blend(vis, np.array([[[0.0, 0.0, 1.0]]]), pred_grad[:, :, None] * 0.5)
        return vis

    def _dens_for_plot(self):
        d = self.dens_img
        if d.shape != tuple(self.img_size):       # histogram bins leave the last row / column out: pad to the image size
            out = np.zeros(self.img_size, dtype=float)
            out[:d.shape[0], :d.shape[1]] = d[:self.img_size[0], :self.img_size[1]]
            d = out
        return d

    def _finish_plot(self, vis):
        import cv2
        vis = (np.clip(vis, 0.0, 1.0) * 255.0).astype(np.uint8)
        for cls, colour in ((0, (255, 128, 0)), (1, (0, 0, 255))):
            for i in np.where(self.sup_y == cls)[0]:
                cv2.circle(vis, (int(self.sup_X_img[i, 1]), int(self.sup_X_img[i, 0])), 5, colour, 2)
        return vis

    def semisup_image_plot(self, pred_y1, pred_grad):
        return self._finish_plot(self._plot_base(pred_y1, pred_grad, 2))


class SplitClassificationDataset2D(ClassificationDataset2D):
    """Draws the supervised subset: class-balanced prefix of a shuffle, or a stratified split (reference :107-127)."""

    def __init__(self, X, y, img_size, n_sup, balance_classes, rng):
        if balance_classes:
            n_classes = y.max() + 1
            per_class = n_sup // n_classes
            sup, unsup = [], []
            for c in range(n_classes):
                idx = np.arange(len(y))[y == c]
                rng.shuffle(idx)
                sup.append(idx[:per_class])
                unsup.append(idx)
            sup_indices, unsup_indices = np.concatenate(sup, axis=0), np.concatenate(unsup, axis=0)
        else:
            from sklearn.model_selection import StratifiedShuffleSplit
            splitter = StratifiedShuffleSplit(n_splits=1, test_size=n_sup, random_state=rng)
            _, sup_indices = next(splitter.split(y, y))
            unsup_indices = np.arange(len(y))
        super(SplitClassificationDataset2D, self).__init__(X, y, img_size, sup_indices, unsup_indices)


class ClassificationDatasetFromImage2D(SplitClassificationDataset2D):
    """Data set sampled from a two-region image; keeps the image and its edge map for the plots (reference :131-177)."""

    def __init__(self, image, X, y, img_size, n_sup, balance_classes, rng):
        self.img_size = img_size
        self.img_scale = np.array(img_size).astype(float)
        super(ClassificationDatasetFromImage2D, self).__init__(X, y, img_size, n_sup, balance_classes, rng)
        self.image = image
        self.image_edges = _roberts(self.image)

    def semisup_image_plot(self, pred_y1, pred_grad):
        vis = self._plot_base(pred_y1, pred_grad, 1)
        vis = blend(vis, np.array([[[1.0, 0.0, 1.0]]]), self.image_edges[:, :, None] * 0.5)
        return self._finish_plot(vis)


def classification_dataset_from_image(image_path, region_erode_radius, img_noise_std, n_sup, balance_classes, rng):
    """Two classes = the two regions of a black / white image, eroded away from the boundary, one sample per remaining pixel
    plus Gaussian position noise (reference :180-209)."""
    from PIL import Image
    from scipy.ndimage import binary_erosion
    img = _grey_float(np.array(Image.open(image_path)))
    img_bin = img >= 0.5
    img_size = img_bin.shape
    if region_erode_radius > 0:
        cls_1 = binary_erosion(img_bin, iterations=region_erode_radius)
        cls_0 = binary_erosion(~img_bin, iterations=region_erode_radius)
    else:
        cls_1, cls_0 = img_bin, ~img_bin
    y0, x0 = np.where(cls_0)
    y1, x1 = np.where(cls_1)
    X_img = np.append(np.stack([y0, x0], axis=1), np.stack([y1, x1], axis=1), axis=0)
    y = np.append(np.zeros((len(y0),), dtype=int), np.ones((len(y1),), dtype=int), axis=0)
    X_img = X_img + rng.normal(loc=0, scale=img_noise_std, size=X_img.shape)
    X_real = (X_img / np.array(img_size)) * 2 - 1
    return ClassificationDatasetFromImage2D(img, X_real, y, img_size, n_sup, balance_classes, rng)


def spiral_classification_dataset(n_sup, balance_classes, rng, N=5000, spiral_radius=20, img_size=(256, 256)):
    """Two interleaved spiral arms, N samples each, uniform in area (sqrt of uniform squared radii) (reference :212-232)."""
    r0 = np.sqrt(rng.uniform(low=1.0, high=spiral_radius ** 2, size=(N,)))
    r1 = np.sqrt(rng.uniform(low=1.0, high=spiral_radius ** 2, size=(N,)))
    radius = np.append(r0, r1, axis=0)
    theta = np.append(r0 * 0.5, r1 * 0.5 + np.pi, axis=0)
    X = np.stack([np.sin(theta) * radius, np.cos(theta) * radius], axis=1)
    y = np.append(np.zeros(r0.shape, dtype=int), np.ones(r1.shape, dtype=int), axis=0)
    X = (X + rng.normal(size=X.shape) * 0.2) / spiral_radius
    return SplitClassificationDataset2D(X, y, img_size, n_sup, balance_classes, rng)


def classification_dataset_from_pickle(path, img_size=(512, 512)):
    """[B200 build] A data set written by the reference's `generate_data.py clf` command (`data/toy2d/curve_mask_v3_35.pkl`:
    keys clf_sup_X / clf_sup_y / clf_unsup_X / clf_unsup_y; the unsupervised set is the whole data set, reference :123)."""
    with open(path, 'rb') as f:
        data = pickle.load(f)
    X, y = np.asarray(data['clf_unsup_X']), np.asarray(data['clf_unsup_y'])
    ds = ClassificationDataset2D(X, y, tuple(img_size), np.arange(0), np.arange(len(y)))
    ds.sup_X, ds.sup_y = np.asarray(data['clf_sup_X']), np.asarray(data['clf_sup_y'])
    ds.sup_X_img = ds.real_to_img(ds.sup_X)
    return ds
