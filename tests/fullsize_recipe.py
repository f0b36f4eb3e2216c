"""Recipe of the FULL-SIZE parity iterations (cfg2 = DeepLab v2, N = 10, 321 x 321, 21 classes; cfg3 = DeepLab v3+,
N = 16, 512 x 512, 19 classes), shared by oracle/gen_golden_fullsize.py (which runs the iteration with the UNMODIFIED reference
modules on the CPU of the development container and commits the results under tests/golden/) and by the GPU parity tests /
bench.py's `parity` block (which rebuild the same inputs from the seeds and run the CUDA path).

Everything is a deterministic function of CPU torch / numpy generators: weights = torch_oracle.synth_state_dict(seed, gain),
batches = cutmix_semisup_seg_b200.synthetic (fixed seeds), CutMix boxes = BoxMaskGenerator with RandomState, the four dropout
keep-masks of the DeepLab v3+ head (one per forward pass, in the reference's order: supervised, teacher view 0, teacher view 1,
student mixed) = Bernoulli(0.5) draws of a seeded CPU generator."""
import torch

CONFIGS = {
    # `gain` scales the final classifier so that the teacher's max-probability straddles the reference's default
    # threshold 0.97 (conf_rate in (0.2, 0.8)); tuned by gen_golden_fullsize.py --tune
    'cfg2': dict(kind='resnet101_deeplab_imagenet', arch='deeplab2', classes=21, n=10, h=321, w=321, lr=3e-5, seed=3,
                 gain=6.0, conf_thresh=0.97, final_key='layer5', iters=3),
    'cfg3': dict(kind='resnet101_deeplabv3plus_imagenet', arch='deeplab3plus', classes=19, n=16, h=512, w=512, lr=1e-5, seed=9,
                 gain=12.0, conf_thresh=0.97, final_key='classifier.classifier.6', iters=3),
    # small stand-ins with the same code path (used to keep the CPU suite fast and to pin the oracle at a size it finishes)
    'cfg3_small': dict(kind='resnet101_deeplabv3plus_imagenet', arch='deeplab3plus', classes=19, n=2, h=256, w=256, lr=1e-5,
                       seed=9, gain=12.0, conf_thresh=0.97, final_key='classifier.classifier.6', iters=3),
}
SUP_SEED, UNSUP_SEED, DROP_SEED = 1000, 2000, 3000


def final_keys(state_dict, cfg):
    return [k for k in state_dict if cfg['final_key'] in k and k.endswith('weight')]


def batches(cfg, mask_generator, compact_masks, it=0):
    """(sup_x, sup_y), unsup dict -- CPU tensors.  compact_masks: 4-int boxes (product) or dense (N,1,H,W) masks (reference)."""
    from cutmix_semisup_seg_b200 import synthetic
    n, h, w, c = cfg['n'], cfg['h'], cfg['w'], cfg['classes']
    sup = synthetic.make_sup_batch(n, h, w, c, SUP_SEED + it)
    uns = synthetic.make_unsup_batch(n, h, w, UNSUP_SEED + it, mask_generator, compact_masks=compact_masks, paired=True)
    return sup, uns


def dropout_masks(cfg, it=0):
    """{'sup','tea0','tea1','stu'} -> NHWC keep-mask (N, H/8, W/8, 256) fp32, or None for architectures without dropout."""
    if cfg['arch'] != 'deeplab3plus':
        return None
    g = torch.Generator().manual_seed(DROP_SEED + it)
    n, fh, fw = cfg['n'], -(-cfg['h'] // 8), -(-cfg['w'] // 8)
    return {k: (torch.rand((n, fh, fw, 256), generator=g) > 0.5).float() for k in ('sup', 'tea0', 'tea1', 'stu')}
