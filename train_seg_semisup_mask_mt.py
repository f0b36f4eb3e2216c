"""Drop-in entry point for the reference's `train_seg_semisup_mask_mt.py`: CutMix / CutOut mean-teacher
semi-supervised segmentation, with the per-iteration hot path (reference lines 287-476) running on the B200
kernels (cutmix_semisup_seg_b200.step.MeanTeacherStep; the outer loop is cutmix_semisup_seg_b200.train_loop.run_training).

The click surface (option names and defaults, reference lines 581-650) and the job function signature are kept.
Differences, all forced by the offline / GPU-native setting:
  * `--dataset synthetic` (new choice) trains on synthetic tensors with the DataLoader's tensor contract;
    `--dataset synthetic_u8` (new choice) trains on seeded uint8 images of assorted sizes that go through the reference's
    train-time transform lists (:147-179: crop / scale / rotation, flips, strong colour on the student's view, normalise)
    ON THE DEVICE (cutmix_semisup_seg_b200.input_pipeline.DeviceTrainPipeline), honouring the split and aug_* options;
    the real datasets need the reference's image archives and decoders (out of scope of the hot path): a usage error says so;
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
    import mask_gen
    from cutmix_semisup_seg_b200 import synthetic, train_loop

    if ':' in mask_prop_range:
        lo, hi = mask_prop_range.split(':')
        mask_prop_range = (float(lo.strip()), float(hi.strip()))
    else:
        mask_prop_range = float(mask_prop_range)
    if mask_mode not in ('zero', 'mix'):
        raise ValueError('Unknown mask_mode {}'.format(mask_mode))
    mask_mix = mask_mode == 'mix'
    mask_generator = mask_gen.BoxMaskGenerator(prop_range=mask_prop_range, n_boxes=boxmask_n_boxes,
                                               random_aspect_ratio=not boxmask_fixed_aspect_ratio,
                                               prop_by_area=not boxmask_by_size, within_bounds=not boxmask_outside_bounds,
                                               invert=not boxmask_no_invert)

    def make_unsup(n, h, w, seed, device):
        return synthetic.make_unsup_batch(n, h, w, seed, mask_generator, mask_mix=mask_mix, paired=aug_strong_colour,
                                          device=device)

    def u8_unsup(batches, n, h, w, seed, device):
        """Unsupervised batch dict from the device pipeline's output (`--dataset synthetic_u8`): with `unsup_paired` the teacher sees
        `sample0` (weak) and the student `sample1` (colour-jittered), reference :313-323; box parameters as for synthetic tensors."""
        import numpy as np
        import torch

        views = train_loop.u8_views
        out = {}
        if mask_mix:
            out['ux0_tea'], out['ux0_stu'], out['um0'] = views(batches[0])
            out['ux1_tea'], out['ux1_stu'], out['um1'] = views(batches[1])
        else:
            out['ux_tea'], out['ux_stu'], out['um'] = views(batches[0])
        boxes = mask_generator.generate_boxes(n, (h, w), rng=np.random.RandomState(12345 + seed))
        out['mask_params'] = torch.from_numpy(boxes).to(device)
        return out

    train_loop.run_training(
        submit_config, settings, make_unsup, mask_generator, mask_mix, u8_unsup=u8_unsup,
        dataset=dataset, model=model, arch=arch, freeze_bn=freeze_bn, opt_type=opt_type, sgd_momentum=sgd_momentum,
        sgd_nesterov=sgd_nesterov, sgd_weight_decay=sgd_weight_decay, learning_rate=learning_rate, lr_sched=lr_sched,
        lr_step_epochs=lr_step_epochs, lr_step_gamma=lr_step_gamma, lr_poly_power=lr_poly_power, teacher_alpha=teacher_alpha,
        bin_fill_holes=bin_fill_holes, crop_size=crop_size, cons_loss_fn=cons_loss_fn, cons_weight=cons_weight,
        conf_thresh=conf_thresh, conf_per_pixel=conf_per_pixel, rampup=rampup, unsup_batch_ratio=unsup_batch_ratio,
        num_epochs=num_epochs, iters_per_epoch=iters_per_epoch, batch_size=batch_size, save_model=save_model,
        no_pretrained=no_pretrained, ddp=ddp, synthetic_classes=synthetic_classes)


@click.command()
@click.option('--job_desc', type=str, default='')
@click.option('--dataset', type=click.Choice(['camvid', 'cityscapes', 'pascal', 'pascal_aug', 'isic2017', 'synthetic', 'synthetic_u8']),
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
