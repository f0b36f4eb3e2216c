"""ORACLE side -- test infrastructure only.  Full-size golden iterations from the UNMODIFIED reference modules.

    cd /tmp && python /root/repo/oracle/gen_golden_fullsize.py cfg3_small cfg2 cfg3     # writes tests/golden/fullsize_<cfg>.npz
    cd /tmp && python /root/repo/oracle/gen_golden_fullsize.py --tune cfg3              # confidence rate vs classifier gain

`iters` consecutive iterations (tests/fullsize_recipe.py) of the reference loop body (train_seg_semisup_mask_mt.py:287-476) at the sizes SURVEY.md 8d names -- cfg2: DeepLab
v2, N = 10, 321 x 321 (the reference script's own batch size); cfg3: DeepLab v3+, N = 16, 512 x 512 -- executed with the
reference's classes imported from /root/reference (networks, EMAWeightOptimizer, BoxMaskGenerator), torch Adam
(foreach=False) on the reference's parameter groups and the loss lines :363-451 restated below as in gen_golden.gen_iteration.
The only intervention: the nn.Dropout(0.5) of torchvision's ASPP is replaced by a module multiplying with a given keep-mask
(x * mask * 2, which is what nn.Dropout computes for that draw), because a random draw cannot be shared between two
implementations (SURVEY.md 8d "DLv3+ dropout: explicit Bernoulli(0.5) masks ... injected into both paths").

Inputs and weights are the seeded recipe of tests/fullsize_recipe.py.  cfg3 takes several minutes per iteration and tens of GB of RAM on 8 cores.
Recorded: the three scalars the reference prints (supervised loss, consistency loss, confidence rate), sub-sampled logits of
the four forward passes, checksums of the post-step student / teacher state."""
import os
import sys
import time
import warnings

REF = os.environ.get('CUTMIX_REF', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT = os.path.join(REPO, 'tests', 'golden')
sys.path.insert(0, REF)                    # reference packages first: architectures, mask_gen, optim_weight_ema
sys.path.insert(1, HERE)
sys.path.append(os.path.join(REPO, 'tests'))
sys.path.append(REPO)                      # cutmix_semisup_seg_b200.synthetic only (no name clash with the reference)
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import mask_gen  # noqa: E402  (reference)
import optim_weight_ema  # noqa: E402  (reference)
from architectures import network_architectures  # noqa: E402  (reference)
import torch_oracle as TO  # noqa: E402
import fullsize_recipe as R  # noqa: E402

assert os.path.realpath(mask_gen.__file__).startswith(os.path.realpath(REF)), 'reference not first on sys.path'
assert os.path.realpath(network_architectures.__file__).startswith(os.path.realpath(REF))


class InjectedDropout(nn.Module):
    """Stands in for nn.Dropout(0.5): y = x * keep * 2 with keep-masks queued by the caller (NCHW)."""

    def __init__(self):
        super(InjectedDropout, self).__init__()
        self.queue = []

    def forward(self, x):
        return x * self.queue.pop(0) * 2.0


def build(cfg):
    net = network_architectures.seg.get(cfg['kind'])(cfg['classes'], pretrained=False)
    sd = TO.synth_state_dict(net.state_dict(), seed=cfg['seed'], logit_gain=cfg['gain'], final_keys=R.final_keys(net.state_dict(), cfg))
    net.load_state_dict(sd)
    drops = []
    for mod in net.modules():
        for name, child in list(mod.named_children()):
            if isinstance(child, nn.Dropout):
                inj = InjectedDropout()
                setattr(mod, name, inj)
                drops.append(inj)
    return net, drops


def sub(t):
    """Sub-sampled logits (every 64th row / column, starting at 8) as float32 numpy."""
    return t.detach()[:, :, 8::64, 8::64].contiguous().numpy()


def run(name):
    cfg = R.CONFIGS[name]
    n, h, w, c, lr, tau = cfg['n'], cfg['h'], cfg['w'], cfg['classes'], cfg['lr'], cfg['conf_thresh']
    t0 = time.time()
    student, sdrop = build(cfg)
    teacher, tdrop = build(cfg)
    for p in teacher.parameters():
        p.requires_grad = False
    optim = torch.optim.Adam([dict(params=student.pretrained_parameters(), lr=lr * 0.1),
                              dict(params=student.new_parameters(), lr=lr)], foreach=False)                 # :90-93
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)                                        # :119-124
    mg = mask_gen.BoxMaskGenerator(prop_range=0.5, n_boxes=1, random_aspect_ratio=True, prop_by_area=True,
                                   within_bounds=True, invert=True)                                          # :131-134
    crit = nn.CrossEntropyLoss(ignore_index=255)                                                             # :126
    rec = dict(sup_loss=[], cons_loss=[], conf_rate=[])
    for it in range(cfg['iters']):
        (sup_x, sup_y), uns = R.batches(cfg, mg, compact_masks=False, it=it)
        dm = R.dropout_masks(cfg, it)
        if dm is not None:
            assert len(sdrop) == 1 and len(tdrop) == 1
            sdrop[0].queue = [dm['sup'].permute(0, 3, 1, 2), dm['stu'].permute(0, 3, 1, 2)]
            tdrop[0].queue = [dm['tea0'].permute(0, 3, 1, 2), dm['tea1'].permute(0, 3, 1, 2)]
        masks = uns['mask_params']
        student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()             # :268-275
        optim.zero_grad()                                                                                    # :290
        logits_sup = student(sup_x)                                                                          # :299
        sup_loss = crit(logits_sup, sup_y[:, 0]); sup_loss.backward()                                        # :300-301
        if it == 0:
            rec['sup_logits'] = sub(logits_sup)
        del logits_sup
        ux_mixed = uns['ux0_stu'] * (1 - masks) + uns['ux1_stu'] * masks                                     # :350
        um_mixed = uns['um0'] * (1 - masks) + uns['um1'] * masks                                             # :351
        with torch.no_grad():
            l0 = teacher(uns['ux0_tea']).detach(); l1 = teacher(uns['ux1_tea']).detach()                     # :354-356
        ls = student(ux_mixed)                                                                               # :358
        lt = l0 * (1 - masks) + l1 * masks                                                                   # :363
        pt = F.softmax(lt, dim=1); ps = F.softmax(ls, dim=1)                                                 # :366-367
        conf = (pt.max(dim=1)[0] >= tau).float()[:, None]                                                    # :409-411
        conf_rate = float(conf.mean())                                                                       # :413
        loss_mask = um_mixed * conf.mean()                                                                   # :415-418
        d = ps - pt
        cons = ((d * d).sum(dim=1, keepdim=True) * loss_mask).mean()                                         # :429-431, :451
        (cons * 1.0).backward()                                                                              # :458-459
        if it == 0:
            rec.update(tea0_logits=sub(l0), tea1_logits=sub(l1), stu_logits=sub(ls))
        del l0, l1, ls, lt, pt, ps, d
        optim.step(); ema.step()                                                                             # :465-467
        rec['sup_loss'].append(float(sup_loss)); rec['cons_loss'].append(float(cons)); rec['conf_rate'].append(conf_rate)
        print('  %s iteration %d: sup_loss %.8f cons_loss %.8e conf_rate %.6f  (%.0f s)' % (name, it, float(sup_loss), float(cons),
                                                                                            conf_rate, time.time() - t0), flush=True)
    tsd, ssd = teacher.state_dict(), student.state_dict()
    first = next(k for k in ssd if k.endswith('conv1.weight'))
    last = R.final_keys(ssd, cfg)[0]
    rec.update(sup_loss=np.array(rec['sup_loss']), cons_loss=np.array(rec['cons_loss']), conf_rate=np.array(rec['conf_rate']),
               teacher_abs_sum=float(sum(v.double().abs().sum() for v in tsd.values() if v.dtype == torch.float32)),
               student_abs_sum=float(sum(v.double().abs().sum() for v in ssd.values() if v.dtype == torch.float32)),
               student_first_sum=float(ssd[first].double().sum()), student_last_sum=float(ssd[last].double().sum()),
               teacher_last_sum=float(tsd[last].double().sum()),
               student_last=ssd[last].detach()[:, :64].numpy().copy(), teacher_last=tsd[last].detach()[:, :64].numpy().copy(),     # first 64 input channels
               bn_running={k: v.numpy().copy() for k, v in ssd.items() if 'classifier.classifier.1.running' in k},
               torch_version=torch.__version__, seconds=time.time() - t0)
    bn = rec.pop('bn_running')
    for k, v in bn.items():
        rec['student_bn__' + k] = v
    np.savez_compressed(os.path.join(OUT, 'fullsize_%s.npz' % name), **rec)
    print(name, 'done in %.0f s' % rec['seconds'], flush=True)


def tune(name):
    """Teacher confidence rate at the reference's default threshold for a few classifier gains (2 images)."""
    base = dict(R.CONFIGS[name]); base['n'] = 2
    for gain in (2.0, 4.0, 6.0, 8.0, 12.0, 20.0):
        cfg = dict(base, gain=gain)
        net, drops = build(cfg)
        net.train(); net.freeze_batchnorm()
        mg = mask_gen.BoxMaskGenerator(prop_range=0.5, n_boxes=1, random_aspect_ratio=True, prop_by_area=True,
                                       within_bounds=True, invert=True)
        (_, _), uns = R.batches(cfg, mg, compact_masks=False)
        dm = R.dropout_masks(cfg)
        if dm is not None:
            drops[0].queue = [dm['tea0'].permute(0, 3, 1, 2)]
        with torch.no_grad():
            p = F.softmax(net(uns['ux0_tea']), dim=1).max(dim=1)[0]
        print(name, 'gain', gain, 'conf_rate@%.2f' % cfg['conf_thresh'], float((p >= cfg['conf_thresh']).float().mean()),
              'mean max-prob', float(p.mean()), flush=True)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    args = sys.argv[1:]
    if args and args[0] == '--tune':
        for a in args[1:]:
            tune(a)
    else:
        for a in args:
            run(a)
