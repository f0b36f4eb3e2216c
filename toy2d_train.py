"""Drop-in for the reference's `toy2d_train.py` (BASELINE config 1, the CPU plumbing configuration): the
2-D toy experiment -- an MLP classifier on points, trained with a supervised loss plus mean-teacher / Pi-model perturbation
consistency -- with the same click surface, job function, printed report and random-number consumption order, so that with the
same torch / numpy seeds it reproduces the reference's numbers (pinned by tests/test_toy2d.py against the UNMODIFIED reference
function, tests/golden/toy2d.json).

What this configuration exercises of the hot path is the wiring around it: `optim_weight_ema.EMAWeightOptimizer` driving a
teacher from a student every iteration, eval / train switching, the loss helpers.  The MLP itself is torch.nn (it is the toy
problem's model, not a segmentation network); on `--device cuda:*` the EMA step is the fused CUDA kernel, on `--device cpu` the
optimiser's host arithmetic (the reference's three roundings).  Differences from the reference:
  * data sets come from `toy2d/generate_data.py` of this repository (no scikit-image / batchup); `--dataset pkl:<path>` (new)
    loads a data set written by the reference's generator, e.g. `data/toy2d/curve_mask_v3_35.pkl`;
  * without `--save_output` the reference opens OpenCV windows; here that needs a display, so a missing display is reported
    instead of crashing inside `cv2.imshow`.
"""
import click

import job_helper


def repeat_forever(sampler):
    """`datapipe.seg_data.RepeatSampler(sampler)` of the reference (infinite repetition of a sampler)."""
    import itertools
    import torch.utils.data

    class _Repeat(torch.utils.data.Sampler):
        def __init__(self, inner):
            self.inner = inner

        def __iter__(self):
            return itertools.chain.from_iterable(itertools.repeat(self.inner))

        def __len__(self):
            return 2 ** 62
    return _Repeat(sampler)


@job_helper.job('toy2d_train', enumerate_job_names=False)
def train_toy2d(submit_config, dataset, region_erode_radius, img_noise_std,
                n_sup, balance_classes, seed,
                sup_path, model, n_hidden, hidden_size, hidden_act, norm_layer,
                perturb_noise_std, dist_contour_range,
                conf_thresh, conf_avg,
                cons_weight, cons_loss_fn, cons_no_dropout,
                learning_rate, teacher_alpha,
                num_epochs, batch_size, render_cons_grad, render_pred, device,
                save_output):
    settings = locals().copy()
    del settings['submit_config']
    import os
    import sys
    import time

    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    import torch.utils.data

    import optim_weight_ema
    from toy2d import generate_data

    print('Command line:')
    print(' '.join(sys.argv))
    print('Settings:')
    print(', '.join(['{}={}'.format(k, settings[k]) for k in sorted(settings.keys())]))

    # ---- data (reference :37-54)
    rng = np.random.RandomState(seed)
    image = None
    if dataset.startswith('img:'):
        ds = generate_data.classification_dataset_from_image(dataset[4:], region_erode_radius, img_noise_std, n_sup,
                                                             balance_classes, rng)
        image = ds.image
    elif dataset.startswith('pkl:'):
        ds = generate_data.classification_dataset_from_pickle(dataset[4:])
    elif dataset == 'spiral':
        ds = generate_data.spiral_classification_dataset(n_sup, balance_classes, rng)
    else:
        print('Unknown dataset {}, should be spiral or img:<path>'.format(dataset))
        return
    if sup_path is not None:
        ds.load_supervised(sup_path)

    # signed distance to the class boundary, for perturbations constrained to its level sets (reference :56-67)
    dist_map = None
    if dist_contour_range > 0.0:
        if image is None:
            print('Constraining perturbations to lying on distance map contours is only supported for \'image\' experiments')
            return
        from scipy.ndimage import distance_transform_edt
        inside = image >= 0.5
        dist_map = distance_transform_edt(inside) * inside - distance_transform_edt(~inside) * (~inside)

    torch_device = torch.device(device)
    try:
        noise_std = np.array([float(v.strip()) for v in perturb_noise_std.split(',')])
    except ValueError:
        noise_std = np.array([6.0, 6.0])
    # the option is in pixels; the network sees [-1, 1] coordinates (reference :77-79)
    noise_std_real = torch.tensor(noise_std / ds.img_scale * 2.0, dtype=torch.float, device=torch_device)

    # ---- model: same module tree and construction order as the reference's Network (:82-120), hence the same state_dict keys
    # and the same draws from the global generator
    class Network(nn.Module):
        def __init__(self):
            super(Network, self).__init__()
            self.drop = nn.Dropout()
            layers, width = [], 2
            for _ in range(n_hidden):
                linear = nn.Linear(width, hidden_size)
                if norm_layer == 'spectral_norm':
                    linear = nn.utils.spectral_norm(linear)
                elif norm_layer == 'weight_norm':
                    linear = nn.utils.weight_norm(linear)
                layers.append(linear)
                if norm_layer == 'batch_norm':
                    layers.append(nn.BatchNorm1d(hidden_size))
                elif norm_layer == 'group_norm':
                    layers.append(nn.GroupNorm(4, hidden_size))
                if hidden_act == 'relu':
                    layers.append(nn.ReLU())
                elif hidden_act == 'lrelu':
                    layers.append(nn.LeakyReLU(0.01))
                else:
                    raise ValueError
                width = hidden_size
            self.hidden = nn.Sequential(*layers)
            self.l_final = nn.Linear(width, 2)

        def forward(self, x, use_dropout=True):
            x = self.hidden(x)
            if use_dropout:
                x = self.drop(x)
            return self.l_final(x)

    student_net = Network().to(torch_device)
    student_optimizer = torch.optim.Adam(list(student_net.parameters()), lr=learning_rate)
    classification_criterion = nn.CrossEntropyLoss()
    if model == 'mean_teacher':
        teacher_net = Network().to(torch_device)
        for p in teacher_net.parameters():
            p.requires_grad = False
        teacher_optimizer = optim_weight_ema.EMAWeightOptimizer(teacher_net, student_net, ema_alpha=teacher_alpha,
                                                                 host_arithmetic=torch_device.type == 'cpu')
        pred_net = teacher_net
    else:
        teacher_net = teacher_optimizer = None
        pred_net = student_net

    def robust_binary_crossentropy(pred, tgt):
        return -(tgt * torch.log(pred + 1.0e-6) + (-tgt + 1.0) * torch.log(-pred + 1.0 + 1e-6))

    t_dist_map = None
    if dist_contour_range > 0.0:
        t_dist_map = torch.tensor(dist_map[None, None, ...], dtype=torch.float, device=torch_device)

    def conf_factor(teacher_prob):
        """Per-sample confidence weight (reference :157-168)."""
        conf = torch.max(teacher_prob, 1)[0].detach()
        if conf_thresh > 0.0:
            fac = (conf >= conf_thresh).float()
        else:
            fac = torch.ones(conf.shape, dtype=torch.float, device=conf.device)
        if conf_avg:
            fac = torch.ones_like(fac) * fac.mean()
        return fac

    def dist_map_weighting(x0, x1):
        """1 where a sample and its perturbed copy lie on (nearly) the same level set of the distance map (reference :174-207)."""
        if t_dist_map is None or dist_contour_range <= 0:
            return torch.ones(len(x0), dtype=torch.float, device=x0.device)
        row0 = torch.cat([x0[:, 1].view(1, 1, -1, 1), x0[:, 0].view(1, 1, -1, 1)], dim=3)
        row1 = torch.cat([x1[:, 1].view(1, 1, -1, 1), x1[:, 0].view(1, 1, -1, 1)], dim=3)
        dist = F.grid_sample(t_dist_map, torch.cat([row0, row1], dim=1))
        delta_sqr = (dist[0, 0, 0, :] - dist[0, 0, 1, :]).pow(2)
        return (delta_sqr <= (dist_contour_range * dist_contour_range)).float()

    def consistency(student_logits, teacher_logits, mod_fac):
        """Per-sample consistency terms (mean over the two classes) times the modulation factor (reference :379-392)."""
        if cons_loss_fn == 'bce':
            per = robust_binary_crossentropy(F.softmax(student_logits, dim=1), F.softmax(teacher_logits, dim=1))
        elif cons_loss_fn == 'var':
            d = F.softmax(student_logits, dim=1) - F.softmax(teacher_logits, dim=1)
            per = d * d
        elif cons_loss_fn == 'logits_var':
            d = student_logits - teacher_logits
            per = d * d
        else:
            raise ValueError
        return per.mean(dim=1) * mod_fac

    # ---- loaders, created in the reference's order (:210-229): each draws from the global generator when iterated
    sup_dataset = torch.utils.data.TensorDataset(torch.tensor(ds.sup_X, dtype=torch.float), torch.tensor(ds.sup_y, dtype=torch.long))
    sup_loader = torch.utils.data.DataLoader(sup_dataset, batch_size, sampler=repeat_forever(torch.utils.data.RandomSampler(sup_dataset)),
                                             num_workers=1)
    unsup_dataset = torch.utils.data.TensorDataset(torch.tensor(ds.unsup_X, dtype=torch.float))
    unsup_loader = torch.utils.data.DataLoader(unsup_dataset, batch_size, sampler=torch.utils.data.RandomSampler(unsup_dataset),
                                               num_workers=1)
    all_loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.tensor(ds.X, dtype=torch.float)), 16384,
                                             shuffle=False, num_workers=1)
    vis_loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.tensor(ds.px_grid_vis, dtype=torch.float)), 16384,
                                             shuffle=False, num_workers=1)

    student_net.eval()
    if teacher_net is not None:
        teacher_net.eval()

    def consistency_grad_magnitude(x0):
        """|d(consistency loss) / d(student logits)| per sample, for the gradient overlay of the plots (reference :237-286)."""
        x1 = x0 + torch.randn(x0.shape, dtype=torch.float, device=torch_device) * noise_std_real[None, :]
        student_optimizer.zero_grad()
        t_logits = teacher_net(x0).detach() if teacher_net is not None else student_net(x0)
        s_logits = student_net(x1)
        mags = [None]
        s_logits.register_hook(lambda g: mags.__setitem__(0, torch.sqrt((g * g).sum(dim=1))))
        mod = conf_factor(F.softmax(t_logits, dim=1)) * dist_map_weighting(x0, x1)
        consistency(s_logits, t_logits, mod).mean().backward()
        return mags[0]

    def render_output_image():
        preds, grads = [], ([] if render_cons_grad else None)
        for (bx,) in vis_loader:
            bx = bx.to(torch_device)
            with torch.no_grad():
                logits = pred_net(bx)
                if render_pred == 'prob':
                    vis = F.softmax(logits, dim=1)[:, 1]
                elif render_pred == 'class':
                    vis = torch.argmax(logits, dim=1)
                else:
                    raise ValueError('Unknown prediction render {}'.format(render_pred))
            preds.append(vis.detach().cpu().numpy())
            if render_cons_grad:
                # (the reference calls this under no_grad and indexes a tuple, so `--render_cons_grad` cannot run there;
                # here the gradient is taken with autograd enabled)
                with torch.enable_grad():
                    grads.append(consistency_grad_magnitude(bx).detach().cpu().numpy())
        preds = np.concatenate(preds, axis=0)
        return ds.semisup_image_plot(preds, np.concatenate(grads, axis=0) if render_cons_grad else None)

    def show(epoch):
        """True if the user asked to stop (ESC in the window)."""
        import cv2
        if save_output and submit_config.run_dir is not None:
            cv2.imwrite(os.path.join(submit_config.run_dir, 'epoch_{:05d}.png'.format(epoch)), render_output_image())
            return False
        if not (os.environ.get('DISPLAY') or os.environ.get('WAYLAND_DISPLAY')):
            raise RuntimeError('no display for the visualisation window: pass --save_output to write epoch_*.png files instead')
        cv2.imshow('Vis', render_output_image())
        return (cv2.waitKey(1) & 255) == 27

    show(0)
    print('|sup|={}'.format(len(ds.sup_X)))
    print('|unsup|={}'.format(len(ds.unsup_X)))
    print('|all|={}'.format(len(ds.X)))
    print('Training...')

    terminated = False
    history = []            # un-rounded per-epoch figures (the report prints six decimals); see `train_toy2d.last_run`
    for epoch in range(num_epochs):
        t1 = time.time()
        student_net.train()
        if teacher_net is not None:
            teacher_net.train()
        sup_acc = conf_acc = cons_acc = n_acc = 0.0
        for (bx, by), (ux,) in zip(sup_loader, unsup_loader):
            bx, by, ux = bx.to(torch_device), by.to(torch_device), ux.to(torch_device)
            ux1 = ux + torch.randn(ux.shape, dtype=torch.float, device=torch_device) * noise_std_real[None, :]
            student_optimizer.zero_grad()
            sup_loss = classification_criterion(student_net(bx), by)
            if cons_weight > 0.0:
                drop = not cons_no_dropout
                if model == 'mean_teacher':
                    t_logits = teacher_net(ux, use_dropout=drop).detach()
                    s_logits = student_net(ux1, use_dropout=drop)
                elif model == 'pi':
                    t_logits = student_net(ux, use_dropout=drop)
                    s_logits = student_net(ux1, use_dropout=drop)
                elif model == 'pi_onebatch':
                    both = student_net(torch.cat([ux, ux1], dim=0), use_dropout=drop)
                    t_logits, s_logits = both[:len(ux)], both[len(ux):]
                else:
                    raise RuntimeError
                weight = dist_map_weighting(ux, ux1)
                conf_fac = conf_factor(F.softmax(t_logits, dim=1))
                cons_loss = consistency(s_logits, t_logits, conf_fac * weight).sum() / weight.sum()
                loss = sup_loss + cons_loss * cons_weight
                conf_rate = float(conf_fac.sum())
            else:
                loss, conf_rate, cons_loss = sup_loss, 0.0, 0.0
            loss.backward()
            student_optimizer.step()
            if teacher_optimizer is not None:
                teacher_optimizer.step()
            sup_acc += float(sup_loss)
            conf_acc += conf_rate
            cons_acc += float(cons_loss)
            n_acc += len(bx)
        if n_acc > 0:
            sup_acc /= n_acc; conf_acc /= n_acc; cons_acc /= n_acc
        student_net.eval()
        if teacher_net is not None:
            teacher_net.eval()
        history.append((sup_acc, conf_acc, cons_acc))
        if show(epoch + 1):
            terminated = True
            break
        t2 = time.time()
        print('Epoch {}: took {:.3f}s: clf loss={:.6f}, conf rate={:.3%}, cons loss={:.6f}'.format(epoch + 1, t2 - t1, sup_acc, conf_acc,
                                                                                                    cons_acc))

    pred_y = []
    with torch.no_grad():
        for (bx,) in all_loader:
            pred_y.append(torch.argmax(pred_net(bx.to(torch_device)), dim=1).detach().cpu().numpy())
    err_rate = (np.concatenate(pred_y, axis=0) != ds.y).mean()
    print('FINAL RESULT: Error rate={:.6%} (supervised and unsupervised samples)'.format(err_rate))
    # for callers that drive the job function directly (tests): the figures behind the report and the trained networks
    train_toy2d.last_run = dict(epochs=history, error_rate=float(err_rate), student_net=student_net, teacher_net=teacher_net)
    if not save_output:
        import cv2
        if not terminated:
            cv2.waitKey()
        cv2.destroyAllWindows()


@click.command()
@click.option('--job_desc', type=str, default='')
@click.option('--dataset', type=str, default='spiral')
@click.option('--region_erode_radius', type=int, default=35)
@click.option('--img_noise_std', type=float, default=2.0)
@click.option('--n_sup', type=int, default=10)
@click.option('--balance_classes', is_flag=True, default=False)
@click.option('--seed', type=int, default=12345)
@click.option('--sup_path', type=click.Path(dir_okay=False, file_okay=True, exists=True))
@click.option('--model', type=click.Choice(['mean_teacher', 'pi', 'pi_onebatch']), default='mean_teacher')
@click.option('--n_hidden', type=int, default=3)
@click.option('--hidden_size', type=int, default=512)
@click.option('--hidden_act', type=click.Choice(['relu', 'lrelu']), default='relu')
@click.option('--norm_layer', type=click.Choice(['none', 'batch_norm', 'weight_norm',
                                                 'spectral_norm', 'group_norm']), default='batch_norm')
@click.option('--perturb_noise_std', type=str, default='6.0')
@click.option('--dist_contour_range', type=float, default=0.0)
@click.option('--conf_thresh', type=float, default=0.97)
@click.option('--conf_avg', is_flag=True, default=False)
@click.option('--cons_weight', type=float, default=10.0)
@click.option('--cons_loss_fn', type=click.Choice(['var', 'bce', 'logits_var']), default='var')
@click.option('--cons_no_dropout', is_flag=True, default=False)
@click.option('--learning_rate', type=float, default=2e-4)
@click.option('--teacher_alpha', type=float, default=0.99)
@click.option('--num_epochs', type=int, default=100)
@click.option('--batch_size', type=int, default=512)
@click.option('--render_cons_grad', is_flag=True, default=False)
@click.option('--render_pred', type=click.Choice(['class', 'prob']), default='prob')
@click.option('--device', type=str, default='cuda:0')
@click.option('--save_output', is_flag=True, default=False)
def experiment(job_desc, dataset, region_erode_radius, img_noise_std, n_sup, balance_classes, seed,
               sup_path, model, n_hidden, hidden_size, hidden_act, norm_layer,
               perturb_noise_std, dist_contour_range,
               conf_thresh, conf_avg,
               cons_weight, cons_loss_fn, cons_no_dropout,
               learning_rate, teacher_alpha,
               num_epochs, batch_size, render_cons_grad, render_pred, device, save_output):
    params = locals().copy()
    train_toy2d.submit(**params)


if __name__ == '__main__':
    experiment()
