"""-m gpu: replaying the iteration from CUDA graphs must be arithmetically identical to launching it eagerly."""
import copy
import os
import sys

import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
import torch_oracle as TO  # noqa: E402
import mask_gen  # noqa: E402
import optim_weight_ema  # noqa: E402
from architectures import network_architectures as na  # noqa: E402

pytestmark = pytest.mark.gpu
dev = torch.device('cuda:0')


def _trainer(kind, classes, sd, use_graph, fused_opt=False):
    from cutmix_semisup_seg_b200 import step as step_mod
    student = na.seg.get(kind)(classes, pretrained=False)
    student.load_state_dict(sd)
    teacher = na.seg.get(kind)(classes, pretrained=False)
    student.to(dev); teacher.to(dev)
    for p in teacher.parameters():
        p.requires_grad = False
    optim = step_mod.make_optimizer(student, 'adam', 1e-5, capturable=True, fused_kernel=fused_opt)
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    for m in list(student.modules()) + list(teacher.modules()):
        if type(m).__name__ == 'B2Dropout':
            m.p = 0.0
    mg = mask_gen.BoxMaskGenerator(0.5, invert=True)
    return step_mod.MeanTeacherStep(student, teacher, optim, ema, mg, conf_thresh=0.5, use_cuda_graph=use_graph), mg


@pytest.mark.parametrize('kind,classes,fused_opt', [('resnet101_deeplab_imagenet', 21, False),
                                                    ('resnet101_deeplabv3plus_imagenet', 19, False),
                                                    ('resnet101_deeplab_imagenet', 21, True)])
def test_graph_replay_equals_eager(kind, classes, fused_opt):
    from cutmix_semisup_seg_b200 import synthetic
    n, h, w = 2, 64, 64
    net = na.seg.get(kind)(classes, pretrained=False)
    final = [k for k in net.state_dict() if ('layer5' in k or 'classifier.classifier.6' in k) and k.endswith('weight')]
    sd = TO.synth_state_dict(net.state_dict(), seed=5, logit_gain=4.0, final_keys=final)
    results = []
    for use_graph in (False, True):
        tr, mg = _trainer(kind, classes, copy.deepcopy(sd), use_graph, fused_opt)
        losses = []
        for it in range(3):
            sup = synthetic.make_sup_batch(n, h, w, classes, 30 + it, device=dev)
            uns = synthetic.make_unsup_batch(n, h, w, 40 + it, mg, device=dev)
            if fused_opt:        # a per-iteration LR schedule must reach the replayed step (pinned-host LR buffer)
                for grp in tr.student_optim.param_groups:
                    grp['lr'] = grp['lr'] * 0.5
            out = tr.step(sup, [uns])
            losses.append([float(out['sup_loss']), float(out['cons_loss']), float(out['conf_rate'])])
        results.append((losses, {k: v.detach().cpu().clone() for k, v in tr.teacher_net.state_dict().items()}))
    (l_eager, t_eager), (l_graph, t_graph) = results
    # the capture's warm-up iteration is rolled back, kernels and their order are the same: bit-identical trajectories
    assert l_eager == l_graph, (l_eager, l_graph)
    assert l_graph[0] != l_graph[1] != l_graph[2]          # replays consume the new inputs
    for k in t_eager:
        assert torch.equal(t_eager[k], t_graph[k]), k


def test_prefetched_inputs_equal_direct_loads():
    """`step(..., prefetch=next_batch)`: the next pinned host batch is copied on a side stream during the iteration and moved
    device-to-device at the start of its own step -- results must be bit-identical to loading every batch directly."""
    from cutmix_semisup_seg_b200 import synthetic
    kind, classes, n, h, w = 'resnet101_deeplab_imagenet', 21, 2, 64, 64
    net = na.seg.get(kind)(classes, pretrained=False)
    final = [k for k in net.state_dict() if 'layer5' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(net.state_dict(), seed=5, logit_gain=4.0, final_keys=final)
    results = []
    for use_prefetch in (False, True):
        tr, mg = _trainer(kind, classes, copy.deepcopy(sd), True)
        batches = [(synthetic.make_sup_batch(n, h, w, classes, 30 + it, pin=True),
                    synthetic.make_unsup_batch(n, h, w, 40 + it, mg, pin=True)) for it in range(4)]
        losses = []
        for it, (sup, uns) in enumerate(batches):
            nxt = (batches[it + 1][0], [batches[it + 1][1]]) if use_prefetch and it + 1 < len(batches) else None
            out = tr.step(sup, [uns], prefetch=nxt)
            assert (tr._staged is not None) == (nxt is not None)         # consumed by this step, re-armed with the next
            losses.append([float(out['sup_loss']), float(out['cons_loss']), float(out['conf_rate'])])
        results.append((losses, {k: v.detach().cpu().clone() for k, v in tr.teacher_net.state_dict().items()}))
    (l_direct, t_direct), (l_pref, t_pref) = results
    assert l_direct == l_pref
    for k, v in t_direct.items():
        assert torch.equal(v, t_pref[k]), k
