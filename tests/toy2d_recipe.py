"""Cases of the toy-2D plumbing configuration (BASELINE config 1) shared by oracle/gen_golden.py::gen_toy2d (which runs the
UNMODIFIED reference `toy2d_train.train_toy2d` on them) and tests/test_toy2d.py (which runs this repository's drop-in)."""
import numpy as np

BASE = dict(dataset='spiral', region_erode_radius=35, img_noise_std=2.0, n_sup=10, balance_classes=False, seed=12345, sup_path=None,
            model='mean_teacher', n_hidden=3, hidden_size=64, hidden_act='relu', norm_layer='batch_norm', perturb_noise_std='6.0',
            dist_contour_range=0.0, conf_thresh=0.97, conf_avg=False, cons_weight=10.0, cons_loss_fn='var', cons_no_dropout=False,
            learning_rate=2e-4, teacher_alpha=0.99, num_epochs=3, batch_size=512, render_cons_grad=False, render_pred='prob',
            device='cpu', save_output=True)

CASES = {
    # mean teacher with BatchNorm (its running statistics are part of the EMA), confidence threshold low enough to fire
    'mean_teacher_bn': dict(conf_thresh=0.55, learning_rate=2e-3),
    # Pi model, no normalisation layer, logits-variance consistency without dropout, averaged confidence, balanced classes
    'pi_plain': dict(model='pi', norm_layer='none', cons_loss_fn='logits_var', cons_no_dropout=True, conf_avg=True, conf_thresh=0.5,
                     balance_classes=True, hidden_act='lrelu', cons_weight=1.0, render_pred='class'),
    # data set from an image, perturbations constrained to the level sets of the boundary distance map, bce consistency
    'image_contours': dict(dataset='img:{mask}', region_erode_radius=6, model='mean_teacher', norm_layer='group_norm',
                           cons_loss_fn='bce', dist_contour_range=4.0, perturb_noise_std='10.0,4.0', conf_thresh=0.0, num_epochs=2,
                           batch_size=256, learning_rate=1e-3),
    'pi_onebatch_supervised_only': dict(model='pi_onebatch', cons_weight=0.0, num_epochs=2),
}
TORCH_SEED = 0


def write_mask_png(path, size=96):
    """A two-region black / white image with a wavy boundary (stand-in for the reference's data/toy2d/curve_mask_v3.png)."""
    from PIL import Image
    yy, xx = np.mgrid[0:size, 0:size]
    boundary = size * 0.5 + size * 0.18 * np.sin(xx * (2.0 * np.pi / size) * 1.5)
    img = (yy > boundary).astype(np.uint8) * 255
    Image.fromarray(np.stack([img, img, img], axis=2)).save(path)


def params(name, mask_path):
    p = dict(BASE)
    p.update(CASES[name])
    p['dataset'] = p['dataset'].format(mask=mask_path)
    return p


def parse_report(text):
    """[(clf loss, conf rate %, cons loss) per epoch], final error rate % -- from the lines the job prints."""
    import re
    epochs = [tuple(float(v) for v in m) for m in
              re.findall(r'Epoch \d+: took [0-9.]+s: clf loss=([-0-9.enaif]+), conf rate=([-0-9.]+)%, cons loss=([-0-9.enaif]+)', text)]
    final = re.search(r'FINAL RESULT: Error rate=([0-9.]+)%', text)
    return epochs, (float(final.group(1)) if final else None)
