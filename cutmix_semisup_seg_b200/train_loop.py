"""The training procedure shared by the drop-in entry points (`train_seg_semisup_mask_mt.py`, `train_seg_semisup_ict.py`,
`train_seg_semisup_aug_mt.py`, `train_seg_semisup_vat_mt.py`):
network / optimiser / EMA construction, LR schedules, the per-epoch loop around `MeanTeacherStep.step`, on-device
evaluation and reporting -- reference train_seg_semisup_mask_mt.py:86-134, 257-530 (the ICT script's outer loop is the
same code, train_seg_semisup_ict.py:62-110, 226-468; only the unsupervised branch of the iteration differs).

Differences from the reference, forced by the offline / GPU-native setting, are listed in the entry points' docstrings.
"""
import os
import time

# Options of the reference's click surface that configure its CPU data pipeline (dataset splits, DataLoader workers, geometric /
# colour augmentation of real images, prediction dumps).  `--dataset synthetic` has no such pipeline: a non-default value cannot
# take effect, which the run says out loud instead of silently accepting it.  {option: reference default}
DATA_PIPELINE_OPTIONS = dict(
    n_sup=100, n_unsup=-1, n_val=-1, split_seed=12345, split_path=None, val_seed=131, save_preds=False, num_workers=4,
    aug_hflip=False, aug_vflip=False, aug_hvflip=False, aug_scale_hung=False, aug_max_scale=1.0, aug_scale_non_uniform=False,
    aug_rot_mag=0.0, aug_colour_brightness=0.4, aug_colour_contrast=0.4, aug_colour_saturation=0.4, aug_colour_hue=0.1,
    aug_colour_prob=0.8, aug_colour_greyscale_prob=0.2, aug_offset_range=16)
NAN_CHECK_EVERY = 32        # iterations between two host reads of the running supervised loss (NaN bail-out)


def check_dataset(dataset, u8_supported=False):
    """Fail before any device / process-group initialisation: `--dataset synthetic` (seeded tensors) and, where the entry point
    supports it, `--dataset synthetic_u8` (seeded uint8 images through the device input pipeline) can train in this build; the real
    data sets need the reference's image archives and decoders, which are outside the B200 hot path (SURVEY.md 8 'out of scope')."""
    import click
    if dataset == 'synthetic_u8' and not u8_supported:
        raise click.UsageError("--dataset 'synthetic_u8' is not wired into this entry point; use --dataset synthetic here.")
    if dataset not in ('synthetic', 'synthetic_u8'):
        raise click.UsageError(
            "--dataset {!r} is not available in the B200 build: the reference's CPU data pipeline (datapipe/, real image archives) "
            "is not part of the hot path.  Use --dataset synthetic (seeded tensors with the DataLoader's tensor contract) or "
            "--dataset synthetic_u8 (seeded uint8 images through the device input pipeline, honours the aug_* options); the "
            "reference CLI default 'pascal_aug' has to be overridden explicitly.".format(dataset))


# data-pipeline options that `--dataset synthetic_u8` consumes (split + every augmentation option of the mask_mt script)
U8_USED_OPTIONS = ('n_sup', 'n_unsup', 'split_seed', 'aug_offset_range', 'aug_hflip', 'aug_vflip', 'aug_hvflip', 'aug_scale_hung', 'aug_max_scale',
                   'aug_scale_non_uniform', 'aug_rot_mag', 'aug_colour_brightness', 'aug_colour_contrast', 'aug_colour_saturation',
                   'aug_colour_hue', 'aug_colour_prob', 'aug_colour_greyscale_prob')


def u8_views(batch):
    """(teacher image, student image, valid mask) of one DeviceTrainPipeline.unsup_batch output: with `unsup_paired` the teacher
    sees `sample0` (weak) and the student `sample1` (colour-jittered), reference train_seg_semisup_mask_mt.py:313-323."""
    if 'sample0' in batch:
        return batch['sample0']['image'], batch['sample1']['image'], batch['sample0']['mask']
    return batch['image'], batch['image'], batch['mask']


def ignored_options(settings, used=()):
    """Names of data-pipeline options given a non-default value although nothing consumes them on synthetic data."""
    out = []
    for key, default in DATA_PIPELINE_OPTIONS.items():
        if key in used or key not in settings:
            continue
        if settings[key] != default:
            out.append(key)
    return out


def run_training(submit_config, settings, make_unsup, mask_generator, mask_mix, *, dataset, model, arch, freeze_bn, opt_type,
                 sgd_momentum, sgd_nesterov, sgd_weight_decay, learning_rate, lr_sched, lr_step_epochs, lr_step_gamma,
                 lr_poly_power, teacher_alpha, bin_fill_holes, crop_size, cons_loss_fn, cons_weight, conf_thresh,
                 conf_per_pixel, rampup, unsup_batch_ratio, num_epochs, iters_per_epoch, batch_size, save_model,
                 no_pretrained, ddp, synthetic_classes, step_options=None, used_options=(), u8_unsup=None, pipeline_options=None, u8_loaders=None):
    """`make_unsup(batch_size, h, w, seed, device)` -> one unsupervised batch dict for MeanTeacherStep.step (CutMix / CutOut
    box parameters, ICT mix factors, augmentation maps or the VAT marker included); `step_options`: extra keyword arguments of
    MeanTeacherStep (VAT radius / direction network); `used_options`: data-pipeline options the calling script does consume
    on synthetic data (e.g. the aug script's rotation / scale magnitudes); `u8_unsup(batches, n, h, w, seed, device)`: builds the
    unsupervised batch dict from `DeviceTrainPipeline.unsup_batch` outputs (`--dataset synthetic_u8`, one per unsupervised loader:
    `u8_loaders`, default two in mix mode); `pipeline_options`: extra DeviceTrainPipeline arguments (the aug script's pair mode)."""
    check_dataset(dataset, u8_supported=u8_unsup is not None)
    if dataset == 'synthetic_u8':
        used_options = tuple(used_options) + U8_USED_OPTIONS
    import numpy as np
    import torch
    from architectures import network_architectures
    import evaluation
    import lr_schedules
    import optim_weight_ema
    from . import step as step_mod, synthetic

    crop = None if crop_size == '' else [int(x.strip()) for x in crop_size.split(',')]

    rank, world = 0, 1
    if ddp:
        import torch.distributed as dist
        dist.init_process_group('nccl')
        rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch_device = torch.device('cuda', local)
    torch.cuda.set_device(torch_device)

    ign = ignored_options(settings, used_options)
    if ign and rank == 0:
        print('WARNING: --dataset synthetic has no data pipeline; these options have no effect: {}'.format(', '.join(sorted(ign))))
    n_classes = synthetic_classes
    if bin_fill_holes and n_classes != 2:
        print('Binary hole filling can only be used with binary (2-class) segmentation datasets')
        return
    if crop is None:
        crop = [321, 321]
    print('Loaded data')

    NetClass = network_architectures.seg.get(arch)
    import inspect
    takes_pretrained = 'pretrained' in inspect.signature(NetClass).parameters      # densenet161unet*(num_classes) do not

    def build_net(pretrained):
        return NetClass(n_classes, pretrained=pretrained) if takes_pretrained else NetClass(n_classes)
    student_net = build_net(not no_pretrained).to(torch_device)
    if ddp:
        # identical replicas must not depend on every rank drawing the same initial weights (RNG state, cached files):
        # rank 0's parameters AND buffers are the model, before the EMA teacher copies them
        with torch.no_grad():
            for t in student_net.state_dict().values():
                dist.broadcast(t, src=0)
    # one fused launch for the optimiser step + the teacher's EMA step (cutmix_semisup_seg_b200/optim.py: torch's per-tensor
    # arithmetic incl. the duplicated DeepLab v2 group); B200SEG_FUSED_OPT=0 keeps torch.optim + the EMA kernel
    student_optim = step_mod.make_optimizer(student_net, opt_type, learning_rate, sgd_momentum, sgd_nesterov, sgd_weight_decay,
                                            fused_kernel=os.environ.get('B200SEG_FUSED_OPT', '1') != '0')
    if model == 'mean_teacher':
        teacher_net = build_net(False).to(torch_device)
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

    if iters_per_epoch == -1:
        iters_per_epoch = 100
    total_iters = iters_per_epoch * num_epochs
    lr_epoch_scheduler, lr_iter_scheduler = lr_schedules.make_lr_schedulers(
        optimizer=student_optim, total_iters=total_iters, schedule_type=lr_sched, step_epochs=lr_step_epochs,
        step_gamma=lr_step_gamma, poly_power=lr_poly_power)

    trainer = step_mod.MeanTeacherStep(student_net, teacher_net, student_optim, teacher_optim, mask_generator,
                                       cons_loss_fn=cons_loss_fn, cons_weight=cons_weight, conf_thresh=conf_thresh,
                                       conf_per_pixel=conf_per_pixel, rampup=rampup, mask_mix=mask_mix,
                                       unsup_batch_ratio=unsup_batch_ratio, dist_group=True if ddp else None,
                                       **(step_options or {}))

    if rank == 0:
        print('Settings:')
        print(', '.join(['{}={}'.format(key, settings[key]) for key in sorted(list(settings.keys()))]))

    h, w = crop
    iter_i = 0
    source = pipe = None
    if dataset == 'synthetic_u8':
        # the reference's train-time data path behind the decoder: semi-supervised split, two endless random samplers, the
        # transform lists of :147-179 on the device (crop / scale / rotate, flips, colour jitter, normalise)
        from . import input_pipeline
        source = synthetic.U8ImageSource(max(4 * batch_size, 32), (h, w), n_classes, 4321 + rank, torch_device,
                                         n_sup=settings['n_sup'], n_unsup=settings['n_unsup'], split_seed=settings['split_seed'])
        pipe = input_pipeline.DeviceTrainPipeline(
            (h, w), student_net.MEAN, student_net.STD, settings['aug_hflip'], settings['aug_vflip'], settings['aug_hvflip'],
            settings['aug_scale_hung'], settings['aug_max_scale'], settings['aug_scale_non_uniform'], settings['aug_rot_mag'],
            settings.get('aug_strong_colour', False), settings['aug_colour_brightness'], settings['aug_colour_contrast'],
            settings['aug_colour_saturation'], settings['aug_colour_hue'], settings['aug_colour_prob'],
            settings['aug_colour_greyscale_prob'], rng=np.random.RandomState(1000 + rank), flip_rng=np.random.RandomState(2000 + rank),
            **(pipeline_options or {}))
        sample_gen = torch.Generator().manual_seed(3000 + rank)
        sup_iter = source.sampler(source.sup_ndx, batch_size, sample_gen)
        unsup_iter = source.sampler(source.unsup_ndx, batch_size, sample_gen)
        if rank == 0:
            print('synthetic_u8: {} images, {} supervised / {} unsupervised; {} geometric stage{}{}'.format(
                len(source), len(source.sup_ndx), len(source.unsup_ndx), pipe.kind, ', flips' if pipe.any_flip else '',
                ', strong colour (paired)' if pipe.unsup_paired else ''))
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
        conf_ramp_acc = 0.0
        n_unsup_batches = 0
        for it in range(iters_per_epoch):
            if lr_iter_scheduler is not None:
                lr_iter_scheduler.step(iter_i)
            seed = (iter_i * world + rank) * 7
            unsup = []
            if source is not None:
                b = pipe.sup_batch(source.sup(next(sup_iter)))
                sup = (b['image'], b['labels'])
                if cons_weight > 0.0:
                    for r in range(unsup_batch_ratio):
                        # one batch per unsupervised loader (:205-213: two loaders over the same sampler in mix mode)
                        batches = [pipe.unsup_batch(source.unsup(next(unsup_iter))) for _ in range(u8_loaders or (2 if mask_mix else 1))]
                        unsup.append(u8_unsup(batches, batch_size, h, w, seed + 1 + r, torch_device))
            else:
                sup = synthetic.make_sup_batch(batch_size, h, w, n_classes, seed, device=torch_device)
                if cons_weight > 0.0:
                    for r in range(unsup_batch_ratio):
                        unsup.append(make_unsup(batch_size, h, w, seed + 1 + r, torch_device))
            out = trainer.step(sup, unsup, ramp_val=ramp_val)
            sup_acc += out['sup_loss']
            if out['cons_loss'] is not None:
                cons_acc += out['cons_loss']
                # reference :406-420: the confidence rate is accumulated per unsupervised batch when thresholding is on;
                # without a threshold the ramp value stands in, but only `elif rampup > 0` -- otherwise nothing is added
                if conf_thresh > 0.0:
                    conf_acc += out['conf_rate']             # (already the sum over this iteration's unsup batches)
                elif rampup > 0:
                    conf_ramp_acc += ramp_val * len(unsup)
                n_unsup_batches += len(unsup)
            iter_i += 1
            # reference :468-471 reads the loss back every iteration to bail out on NaN; here the running sum is read every
            # NAN_CHECK_EVERY iterations (a NaN term makes the sum NaN), so a dead network stops within that many iterations
            if (it + 1) % NAN_CHECK_EVERY == 0 and it + 1 < iters_per_epoch and bool(torch.isnan(sup_acc)):
                print('NaN detected; network dead, bailing.')
                return
        sup_loss_val = float(sup_acc) / iters_per_epoch
        if np.isnan(sup_loss_val):
            print('NaN detected; network dead, bailing.')
            return
        cons_val = float(cons_acc) / max(n_unsup_batches, 1)
        conf_val = (float(conf_acc) + conf_ramp_acc) / max(n_unsup_batches, 1)

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
