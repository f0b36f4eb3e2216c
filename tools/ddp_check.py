"""torchrun --nproc-per-node 2 tools/ddp_check.py : data-parallel check of the bucketed, overlapped gradient all-reduce on real
GPUs (NCCL).  Every rank runs the same small DeepLab v3+ CutMix iterations on ITS OWN batches, in CUDA-graph mode, once with the
bucket all-reduces launched between graph segments (B200SEG_GRAD_BUCKETS=4) and once with the single all-reduce after the pass
(=1).  Checks: replicas stay bit-identical across ranks, and the two schedules produce the same parameters (same averages)."""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(buckets, rank, dev):
    os.environ['B200SEG_GRAD_BUCKETS'] = str(buckets)
    from architectures import network_architectures as na
    import mask_gen, optim_weight_ema
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    torch.manual_seed(0)
    kind, c, n, h, w = 'resnet101_deeplabv3plus_imagenet', 19, 2, 128, 128
    student = na.seg.get(kind)(c, pretrained=False)
    sd = synthetic.synth_state_dict(student.state_dict(), seed=5, logit_gain=4.0,
                                    final_keys=[k for k in student.state_dict() if 'classifier.classifier.6' in k and k.endswith('weight')])
    student.load_state_dict(sd)
    teacher = na.seg.get(kind)(c, pretrained=False)
    student.to(dev); teacher.to(dev)
    for p in teacher.parameters():
        p.requires_grad = False
    optim = step_mod.make_optimizer(student, 'adam', 1e-5, fused_kernel=True)
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    mg = mask_gen.BoxMaskGenerator(0.5, invert=True)
    tr = step_mod.MeanTeacherStep(student, teacher, optim, ema, mg, conf_thresh=0.5, dist_group=True, use_cuda_graph=True)
    losses = []
    for it in range(3):
        sup = synthetic.make_sup_batch(n, h, w, c, 100 + 10 * rank + it, device=dev)
        uns = synthetic.make_unsup_batch(n, h, w, 200 + 10 * rank + it, mg, device=dev)
        out = tr.step(sup, [uns])
        losses.append(float(out['sup_loss']))
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().reshape(-1) for p in student.parameters()])
    segs = len(tr._graph[0])
    return flat, losses, segs, (tr._bucketed.launched if tr._bucketed is not None else 0)


def main():
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    world = dist.get_world_size()
    res = {}
    for buckets in (4, 1):
        flat, losses, segs, launched = run(buckets, rank, dev)
        others = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(others, flat)
        same = all(bool(torch.equal(others[0], o)) for o in others[1:])
        res[buckets] = (flat, same, losses, segs, launched)
        if rank == 0:
            print('buckets={}: graph segments {} bucket all-reduces launched {} replicas identical {} losses {}'.format(
                buckets, segs, launched, same, ['%.6f' % v for v in losses]), flush=True)
    diff = (res[4][0] - res[1][0]).abs().max().item() / (res[1][0].abs().max().item() + 1e-30)
    if rank == 0:
        print('overlapped buckets vs single all-reduce: max rel parameter difference {:.3e}'.format(diff), flush=True)
        ok = res[4][1] and res[1][1] and diff < 1e-5 and res[4][3] > 1 and res[4][4] > 0
        print('DDP_CHECK_OK' if ok else 'DDP_CHECK_FAILED', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
