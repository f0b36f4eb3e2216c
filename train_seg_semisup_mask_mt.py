"""Drop-in entry point for the reference's `train_seg_semisup_mask_mt.py`: CutMix / CutOut mean-teacher
semi-supervised segmentation, with the per-iteration hot path (reference lines 287-476) running on the B200
kernels (cutmix_semisup_seg_b200.step.MeanTeacherStep).

The click surface (option names and defaults, reference lines 581-650) and the job function signature are kept.
Differences, all forced by the offline / GPU-native setting:
  * `--dataset synthetic` (new choice) trains on synthetic tensors with the DataLoader's tensor contract; the real
    datasets need the reference's CPU data pipeline (`datapipe`, out of scope of the hot path): if that package is
    importable it is used unchanged, otherwise a clear error is raised;
  * `--arch` networks are built with `pretrained` only if the weights are cached locally (`--no_pretrained`);
  * losses / confidence rate are kept on the device and read once per epoch (the reference synchronises three
    times per iteration, lines 413, 461, 469); the NaN bail-out is checked at the same point;
  * `--ddp` (new): data parallelism, one process per GPU launched by torchrun; gradients are averaged with one
    all-reduce per iteration.
"""
import click

import job_helper


@job_helper.job('train_seg_semisup_mask_mt', enumerate_job_names=False)
def train_seg_semisup_mask_mt(submit_config, dataset, model, arch, freeze_bn,
                              opt_type, sgd_momentum, sgd_nesterov, sgd_weight_decay,
                              learning_rate, lr_sched, lr_step_epochs, lr_step_gamma, lr_poly_power,
                              teacher_alpha, bin_fill_holes,
                              crop_size, aug_hflip, aug_vflip, aug_hvflip, aug_scale_hung, aug_max_scale,
                              aug_scale_non_uniform, aug_rot_mag,
                              aug_strong_colour, aug_colour_brightness, aug_colour_contrast, aug_colour_saturation,
                              aug_colour_hue, aug_colour_prob, aug_colour_greyscale_prob,
                              mask_mode, mask_prop_range,
                              boxmask_n_boxes, boxmask_fixed_aspect_ratio, boxmask_by_size, boxmask_outside_bounds,
                              boxmask_no_invert,
                              cons_loss_fn, cons_weight, conf_thresh, conf_per_pixel, rampup, unsup_batch_ratio,
                              num_epochs, iters_per_epoch, batch_size,
                              n_sup, n_unsup, n_val, split_seed, split_path, val_seed, save_preds, save_model,
                              num_workers, no_pretrained=False, ddp=False, synthetic_classes=21):
    settings = locals().copy()
    del settings['submit_config']
    import os
    import time
    import warnings
    import numpy as np
    import torch
    from architectures import network_architectures
    import evaluation
    import lr_schedules
    import mask_gen
    import optim_weight_ema
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic

    if ':' in mask_prop_range:
        lo, hi = mask_prop_range.split(':')
        mask_prop_range = (float(lo.strip()), float(hi.strip()))
    else:
        mask_prop_range = float(mask_prop_range)
    if mask_mode not in ('zero', 'mix'):
        raise ValueError('Unknown mask_mode {}'.format(mask_mode))
    mask_mix = mask_mode == 'mix'
    crop = None if crop_size == '' else [int(x.strip()) for x in crop_size.split(',')]

    rank, world = 0, 1
    if ddp:
        import torch.distributed as dist
        dist.init_process_group('nccl')
        rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch_device = torch.device('cuda', local)
    torch.cuda.set_device(torch_device)

    if dataset != 'synthetic':
        try:
            from datapipe import datasets  # noqa: F401  (the reference's CPU data pipeline, if the user provides it)
        except ImportError:
            raise NotImplementedError(
                'dataset {!r} needs the reference data pipeline (datapipe/, CPU, out of scope of the B200 hot path); put '
                'the reference repository on PYTHONPATH or use --dataset synthetic'.format(dataset))
        raise NotImplementedError('real-dataset loaders are wired through the reference datapipe in a later round; '
                                  'use --dataset synthetic')
    n_classes = synthetic_classes
    if bin_fill_holes and n_classes != 2:
        print('Binary hole filling can only be used with binary (2-class) segmentation datasets')
        return
    if crop is None:
        crop = [321, 321]
    print('Loaded data')

    NetClass = network_architectures.seg.get(arch)
    student_net = NetClass(n_classes, pretrained=not no_pretrained).to(torch_device)
    # one fused launch for the optimiser step + the teacher's EMA step (cutmix_semisup_seg_b200/optim.py: torch's per-tensor
    # arithmetic incl. the duplicated DeepLab v2 group); B200SEG_FUSED_OPT=0 keeps torch.optim + the EMA kernel
    student_optim = step_mod.make_optimizer(student_net, opt_type, learning_rate, sgd_momentum, sgd_nesterov, sgd_weight_decay,
                                            fused_kernel=os.environ.get('B200SEG_FUSED_OPT', '1') != '0')
    if model == 'mean_teacher':
        teacher_net = NetClass(n_classes, pretrained=False).to(torch_device)
        for p in teacher_net.parameters():
            p.requires_grad = False
        teacher_optim = optim_weight_ema.EMAWeightOptimizer(teacher_net, student_net, teacher_alpha)
        eval_net = teacher_net
    elif model == 'pi':
        teacher_net, teacher_optim, eval_net = student_net, None, student_net
    else:
        print('Unknown model type {}'.format(model))
        return
    if freeze_bn and not hasattr(student_net, 'freeze_batchnorm'):
        raise ValueError('Network {} does not support batchnorm freezing'.format(arch))
    print('Built network')

    mask_generator = mask_gen.BoxMaskGenerator(prop_range=mask_prop_range, n_boxes=boxmask_n_boxes,
                                               random_aspect_ratio=not boxmask_fixed_aspect_ratio,
                                               prop_by_area=not boxmask_by_size, within_bounds=not boxmask_outside_bounds,
                                               invert=not boxmask_no_invert)
    if iters_per_epoch == -1:
        iters_per_epoch = 100
    total_iters = iters_per_epoch * num_epochs
    lr_epoch_scheduler, lr_iter_scheduler = lr_schedules.make_lr_schedulers(
        optimizer=student_optim, total_iters=total_iters, schedule_type=lr_sched, step_epochs=lr_step_epochs,
        step_gamma=lr_step_gamma, poly_power=lr_poly_power)

    trainer = step_mod.MeanTeacherStep(student_net, teacher_net, student_optim, teacher_optim, mask_generator,
                                       cons_loss_fn=cons_loss_fn, cons_weight=cons_weight, conf_thresh=conf_thresh,
                                       conf_per_pixel=conf_per_pixel, rampup=rampup, mask_mix=mask_mix,
                                       unsup_batch_ratio=unsup_batch_ratio, dist_group=True if ddp else None)

    if rank == 0:
        print('Settings:')
        print(', '.join(['{}={}'.format(key, settings[key]) for key in sorted(list(settings.keys()))]))

    h, w = crop
    iter_i = 0
    print('Training...')
    for epoch_i in range(num_epochs):
        if lr_epoch_scheduler is not None:
            lr_epoch_scheduler.step(epoch_i)
        t1 = time.time()
        ramp_val = network_architectures.sigmoid_rampup(epoch_i, rampup) if rampup > 0 else 1.0
        student_net.train()
        if teacher_net is not student_net:
            teacher_net.train()
        if freeze_bn:
            student_net.freeze_batchnorm()
            if teacher_net is not student_net:
                teacher_net.freeze_batchnorm()
        sup_acc = torch.zeros((), device=torch_device)
        cons_acc = torch.zeros((), device=torch_device)
        conf_acc = torch.zeros((), device=torch_device)
        n_unsup_batches = 0
        for it in range(iters_per_epoch):
            if lr_iter_scheduler is not None:
                lr_iter_scheduler.step(iter_i)
            seed = (iter_i * world + rank) * 7
            sup = synthetic.make_sup_batch(batch_size, h, w, n_classes, seed, device=torch_device)
            unsup = []
            if cons_weight > 0.0:
                for r in range(unsup_batch_ratio):
                    unsup.append(synthetic.make_unsup_batch(batch_size, h, w, seed + 1 + r, mask_generator, mask_mix=mask_mix,
                                                            paired=aug_strong_colour, device=torch_device))
            out = trainer.step(sup, unsup, ramp_val=ramp_val)
            sup_acc += out['sup_loss']
            if out['cons_loss'] is not None:
                cons_acc += out['cons_loss']
                conf_acc += out['conf_rate'] if conf_thresh > 0.0 else ramp_val
                n_unsup_batches += len(unsup)
            iter_i += 1
        sup_loss_val = float(sup_acc) / iters_per_epoch                # the only host sync of the epoch
        if np.isnan(sup_loss_val):
            print('NaN detected; network dead, bailing.')
            return
        cons_val = float(cons_acc) / max(n_unsup_batches, 1)
        conf_val = float(conf_acc) / max(n_unsup_batches, 1)

        eval_net.eval()
        iou_eval = evaluation.EvaluatorIoU(n_classes, bin_fill_holes)
        with torch.no_grad():
            vx, vy = synthetic.make_sup_batch(min(batch_size, 4), h, w, n_classes, 999, device=torch_device)
            if bin_fill_holes:      # hole filling is a CPU (scipy) post-process of the argmax map, as in the reference
                pred = torch.argmax(eval_net(vx), dim=1).cpu().numpy()
                truth = vy.cpu().numpy()
                for i in range(len(pred)):
                    iou_eval.sample(truth[i, 0], pred[i], ignore_value=255)
            else:                   # fused argmax + confusion matrix on the device, one C*C read-back per epoch
                iou_eval.sample_logits(eval_net(vx), vy, ignore_value=255)
        iou = iou_eval.score()
        t2 = time.time()
        if rank == 0:
            print('Epoch {}: took {:.3f}s, TRAIN clf loss={:.6f}, consistency loss={:.6f}, conf rate={:.3%}, VAL mIoU={:.3%}, '
                  '{:.1f} images/s'.format(epoch_i + 1, t2 - t1, sup_loss_val, cons_val, conf_val, iou.mean(),
                                           iters_per_epoch * batch_size * world / (t2 - t1)))
            print('-- {}'.format(', '.join(['{:.3%}'.format(x) for x in iou])))

    if save_model and rank == 0:
        torch.save(eval_net, os.path.join(submit_config.run_dir, 'model.pth'))
    if ddp:
        dist.destroy_process_group()


@click.command()
@click.option('--job_desc', type=str, default='')
@click.option('--dataset', type=click.Choice(['camvid', 'cityscapes', 'pascal', 'pascal_aug', 'isic2017', 'synthetic']),
              default='pascal_aug')
@click.option('--model', type=click.Choice(['mean_teacher', 'pi']), default='mean_teacher')
@click.option('--arch', type=str, default='resnet101_deeplab_imagenet')
@click.option('--freeze_bn', is_flag=True, default=False)
@click.option('--opt_type', type=click.Choice(['adam', 'sgd']), default='adam')
@click.option('--sgd_momentum', type=float, default=0.9)
@click.option('--sgd_nesterov', is_flag=True, default=False)
@click.option('--sgd_weight_decay', type=float, default=5e-4)
@click.option('--learning_rate', type=float, default=1e-4)
@click.option('--lr_sched', type=click.Choice(['none', 'stepped', 'cosine', 'poly']), default='none')
@click.option('--lr_step_epochs', type=str, default='')
@click.option('--lr_step_gamma', type=float, default=0.1)
@click.option('--lr_poly_power', type=float, default=0.9)
@click.option('--teacher_alpha', type=float, default=0.99)
@click.option('--bin_fill_holes', is_flag=True, default=False)
@click.option('--crop_size', type=str, default='321,321')
@click.option('--aug_hflip', is_flag=True, default=False)
@click.option('--aug_vflip', is_flag=True, default=False)
@click.option('--aug_hvflip', is_flag=True, default=False)
@click.option('--aug_scale_hung', is_flag=True, default=False)
@click.option('--aug_max_scale', type=float, default=1.0)
@click.option('--aug_scale_non_uniform', is_flag=True, default=False)
@click.option('--aug_rot_mag', type=float, default=0.0)
@click.option('--aug_strong_colour', is_flag=True, default=False)
@click.option('--aug_colour_brightness', type=float, default=0.4)
@click.option('--aug_colour_contrast', type=float, default=0.4)
@click.option('--aug_colour_saturation', type=float, default=0.4)
@click.option('--aug_colour_hue', type=float, default=0.1)
@click.option('--aug_colour_prob', type=float, default=0.8)
@click.option('--aug_colour_greyscale_prob', type=float, default=0.2)
@click.option('--mask_mode', type=click.Choice(['zero', 'mix']), default='mix')
@click.option('--mask_prop_range', type=str, default='0.5')
@click.option('--boxmask_n_boxes', type=int, default=1)
@click.option('--boxmask_fixed_aspect_ratio', is_flag=True, default=False)
@click.option('--boxmask_by_size', is_flag=True, default=False)
@click.option('--boxmask_outside_bounds', is_flag=True, default=False)
@click.option('--boxmask_no_invert', is_flag=True, default=False)
@click.option('--cons_loss_fn', type=click.Choice(['var', 'bce', 'kld', 'logits_var', 'logits_smoothl1']), default='var')
@click.option('--cons_weight', type=float, default=1.0)
@click.option('--conf_thresh', type=float, default=0.97)
@click.option('--conf_per_pixel', is_flag=True, default=False)
@click.option('--rampup', type=int, default=-1)
@click.option('--unsup_batch_ratio', type=int, default=1)
@click.option('--num_epochs', type=int, default=300)
@click.option('--iters_per_epoch', type=int, default=-1)
@click.option('--batch_size', type=int, default=10)
@click.option('--n_sup', type=int, default=100)
@click.option('--n_unsup', type=int, default=-1)
@click.option('--n_val', type=int, default=-1)
@click.option('--split_seed', type=int, default=12345)
@click.option('--split_path', type=click.Path(readable=True, exists=True))
@click.option('--val_seed', type=int, default=131)
@click.option('--save_preds', is_flag=True, default=False)
@click.option('--save_model', is_flag=True, default=False)
@click.option('--num_workers', type=int, default=4)
@click.option('--no_pretrained', is_flag=True, default=False, help='[B200 build] random init instead of cached weights')
@click.option('--ddp', is_flag=True, default=False, help='[B200 build] data parallel under torchrun (one process per GPU)')
@click.option('--synthetic_classes', type=int, default=21, help='[B200 build] class count of --dataset synthetic')
def experiment(job_desc, dataset, model, arch, freeze_bn,
               opt_type, sgd_momentum, sgd_nesterov, sgd_weight_decay,
               learning_rate, lr_sched, lr_step_epochs, lr_step_gamma, lr_poly_power,
               teacher_alpha, bin_fill_holes,
               crop_size, aug_hflip, aug_vflip, aug_hvflip, aug_scale_hung, aug_max_scale, aug_scale_non_uniform, aug_rot_mag,
               aug_strong_colour, aug_colour_brightness, aug_colour_contrast, aug_colour_saturation, aug_colour_hue,
               aug_colour_prob, aug_colour_greyscale_prob,
               mask_mode, mask_prop_range,
               boxmask_n_boxes, boxmask_fixed_aspect_ratio, boxmask_by_size, boxmask_outside_bounds, boxmask_no_invert,
               cons_loss_fn, cons_weight, conf_thresh, conf_per_pixel, rampup, unsup_batch_ratio,
               num_epochs, iters_per_epoch, batch_size,
               n_sup, n_unsup, n_val, split_seed, split_path, val_seed, save_preds, save_model, num_workers,
               no_pretrained, ddp, synthetic_classes):
    params = locals().copy()
    train_seg_semisup_mask_mt.submit(**params)


if __name__ == '__main__':
    experiment()
