"""Drop-in entry point for the reference's `train_seg_semisup_aug_mt.py` (augmentation-driven consistency, SURVEY.md 8f
row 3): mean-teacher semi-supervised segmentation where teacher and student see two differently augmented crops of the same
unlabelled image and the teacher's prediction is resampled into the student's frame with the affine map between the two
crops (`xf0_to_1`, reference lines 281, 302-312) before the consistency loss (lines 341-398).  The iteration runs on the B200
kernels (cutmix_semisup_seg_b200.step.MeanTeacherStep.unsupervised_aug: the fused augmentation-consistency kernel does the
sampling, both soft-maxes, the confidence mask, the loss and the student gradient in one pass); the outer loop is
cutmix_semisup_seg_b200.train_loop.run_training, shared with train_seg_semisup_mask_mt.py.

The click surface (option names and defaults, reference lines 515-577) and the job function signature are kept; the additions
(`--dataset synthetic`, `--no_pretrained`, `--ddp`, `--synthetic_classes`) are those of train_seg_semisup_mask_mt.py.  With
`--dataset synthetic` the affine maps are drawn from the script's own augmentation ranges (`--aug_rot_mag` degrees,
`--aug_max_scale`, `--aug_offset_range` pixels; one draw per view, like the reference's per-crop transforms).  With
`--dataset synthetic_u8` uint8 images of assorted sizes go through the script's own transform lists (:126-163: pairs of
differently cropped / scaled / rotated / flipped views, colour jitter on the student's view, `xf0_to_1` from the two crops'
matrices) ON THE DEVICE (cutmix_semisup_seg_b200.input_pipeline.DeviceTrainPipeline(script='aug_mt')).
`--cons_loss_fn logits_var` fails on the first unsupervised batch exactly like the reference (its line 373 reads a variable
only the `var` branch assigns).
"""
import click

import job_helper


@job_helper.job('train_seg_semisup_aug_mt', enumerate_job_names=False)
def train_seg_semisup_aug_mt(submit_config, dataset, model, arch, freeze_bn,
                             opt_type, sgd_momentum, sgd_nesterov, sgd_weight_decay,
                             learning_rate, lr_sched, lr_step_epochs, lr_step_gamma, lr_poly_power,
                             teacher_alpha, bin_fill_holes,
                             crop_size, aug_offset_range, aug_hflip, aug_vflip, aug_hvflip,
                             aug_scale_hung, aug_max_scale, aug_scale_non_uniform, aug_rot_mag, aug_free_scale_rot,
                             aug_strong_colour, aug_colour_brightness, aug_colour_contrast, aug_colour_saturation, aug_colour_hue,
                             aug_colour_prob, aug_colour_greyscale_prob,
                             cons_loss_fn, cons_weight, conf_thresh, conf_per_pixel, rampup, unsup_batch_ratio,
                             num_epochs, iters_per_epoch, batch_size,
                             n_sup, n_unsup, n_val, split_seed, split_path, val_seed, save_preds, save_model, num_workers,
                             no_pretrained=False, ddp=False, synthetic_classes=21):
    settings = locals().copy()
    del settings['submit_config']
    from cutmix_semisup_seg_b200 import synthetic, train_loop

    def make_unsup(n, h, w, seed, device):
        return synthetic.make_aug_batch(n, h, w, seed, paired=aug_strong_colour, rot_mag=aug_rot_mag, max_scale=aug_max_scale,
                                        offset_range=aug_offset_range, device=device)

    def u8_unsup(batches, n, h, w, seed, device):
        """`--dataset synthetic_u8`: the pair batch of the device pipeline (script='aug_mt') in MeanTeacherStep's format -- teacher
        view `sample0`, student view `sample1` (colour-jittered with --aug_strong_colour), `xf0_to_1` (reference :291-300)."""
        b = batches[0]
        return dict(ux0=b['sample0']['image'], um0=b['sample0']['mask'], ux1=b['sample1']['image'], um1=b['sample1']['mask'],
                    xf0_to_1=b['xf0_to_1'])

    train_loop.run_training(
        submit_config, settings, make_unsup, None, True, u8_unsup=u8_unsup, u8_loaders=1,
        pipeline_options=dict(script='aug_mt', aug_offset_range=aug_offset_range, aug_free_scale_rot=aug_free_scale_rot),
        dataset=dataset, model=model, arch=arch, freeze_bn=freeze_bn, opt_type=opt_type, sgd_momentum=sgd_momentum,
        sgd_nesterov=sgd_nesterov, sgd_weight_decay=sgd_weight_decay, learning_rate=learning_rate, lr_sched=lr_sched,
        lr_step_epochs=lr_step_epochs, lr_step_gamma=lr_step_gamma, lr_poly_power=lr_poly_power, teacher_alpha=teacher_alpha,
        bin_fill_holes=bin_fill_holes, crop_size=crop_size, cons_loss_fn=cons_loss_fn, cons_weight=cons_weight,
        conf_thresh=conf_thresh, conf_per_pixel=conf_per_pixel, rampup=rampup, unsup_batch_ratio=unsup_batch_ratio,
        num_epochs=num_epochs, iters_per_epoch=iters_per_epoch, batch_size=batch_size, save_model=save_model,
        no_pretrained=no_pretrained, ddp=ddp, synthetic_classes=synthetic_classes,
        used_options=('aug_rot_mag', 'aug_max_scale', 'aug_offset_range'))


@click.command()
@click.option('--job_desc', type=str, default='')
@click.option('--dataset', type=click.Choice(['camvid', 'cityscapes', 'pascal', 'pascal_aug', 'isic2017', 'synthetic', 'synthetic_u8']),
              default='pascal_aug')
@click.option('--model', type=click.Choice(['mean_teacher', 'pi']), default='mean_teacher')
@click.option('--arch', type=str, default='resnet101_deeplab_imagenet')
@click.option('--freeze_bn', is_flag=True, default=False)
@click.option('--opt_type', type=click.Choice(['adam', 'sgd']), default='adam')
@click.option('--sgd_momentum', type=float, default=0.9)
@click.option('--sgd_nesterov', is_flag=True, default=True)
@click.option('--sgd_weight_decay', type=float, default=5e-4)
@click.option('--learning_rate', type=float, default=1e-4)
@click.option('--lr_sched', type=click.Choice(['none', 'stepped', 'cosine', 'poly']), default='none')
@click.option('--lr_step_epochs', type=str, default='')
@click.option('--lr_step_gamma', type=float, default=0.1)
@click.option('--lr_poly_power', type=float, default=0.9)
@click.option('--teacher_alpha', type=float, default=0.99)
@click.option('--bin_fill_holes', is_flag=True, default=False)
@click.option('--crop_size', type=str, default='321,321')
@click.option('--aug_offset_range', type=float, default=16.0)
@click.option('--aug_hflip', is_flag=True, default=False)
@click.option('--aug_vflip', is_flag=True, default=False)
@click.option('--aug_hvflip', is_flag=True, default=False)
@click.option('--aug_scale_hung', is_flag=True, default=False)
@click.option('--aug_max_scale', type=float, default=1.0)
@click.option('--aug_scale_non_uniform', is_flag=True, default=False)
@click.option('--aug_rot_mag', type=float, default=0.0)
@click.option('--aug_free_scale_rot', is_flag=True, default=False)
@click.option('--aug_strong_colour', is_flag=True, default=False)
@click.option('--aug_colour_brightness', type=float, default=0.4)
@click.option('--aug_colour_contrast', type=float, default=0.4)
@click.option('--aug_colour_saturation', type=float, default=0.4)
@click.option('--aug_colour_hue', type=float, default=0.1)
@click.option('--aug_colour_prob', type=float, default=0.8)
@click.option('--aug_colour_greyscale_prob', type=float, default=0.2)
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
               crop_size, aug_offset_range, aug_hflip, aug_vflip, aug_hvflip,
               aug_scale_hung, aug_max_scale, aug_scale_non_uniform, aug_rot_mag, aug_free_scale_rot,
               aug_strong_colour, aug_colour_brightness, aug_colour_contrast, aug_colour_saturation, aug_colour_hue,
               aug_colour_prob, aug_colour_greyscale_prob,
               cons_loss_fn, cons_weight, conf_thresh, conf_per_pixel, rampup, unsup_batch_ratio,
               num_epochs, iters_per_epoch, batch_size,
               n_sup, n_unsup, n_val, split_seed, split_path, val_seed, save_preds, save_model, num_workers,
               no_pretrained, ddp, synthetic_classes):
    params = locals().copy()
    train_seg_semisup_aug_mt.submit(**params)


if __name__ == '__main__':
    experiment()
