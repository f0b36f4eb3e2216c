"""Learning-rate schedules of the reference (lr_schedules.py:4-64): polynomial decay per iteration, plus the
factory returning (epoch_scheduler, iteration_scheduler) for 'none' | 'stepped' | 'cosine' | 'poly'."""
import ast

import torch


class PolynomialLR(torch.optim.lr_scheduler._LRScheduler):
    """lr = base_lr * max((1 - min(max(t / T_max, 0), 1)) ** power, eta_min)."""

    def __init__(self, optimizer, T_max, power=0.9, eta_min=0.0, last_epoch=-1):
        self.T_max, self.power, self.eta_min = T_max, power, eta_min
        super(PolynomialLR, self).__init__(optimizer, last_epoch)

    def get_lr(self):
        if self.last_epoch == 0:
            return self.base_lrs
        progress = min(max(float(self.last_epoch) / float(self.T_max), 0), 1)
        fac = max((1.0 - progress) ** self.power, self.eta_min)
        return [base_lr * fac for base_lr in self.base_lrs]


def make_lr_schedulers(optimizer, total_iters, schedule_type, step_epochs, step_gamma, poly_power=0.9):
    epoch_sched = iter_sched = None
    if schedule_type == 'none':
        pass
    elif schedule_type == 'stepped' and step_epochs is not None and step_epochs.strip() != '':
        if isinstance(step_epochs, str):
            step_epochs = ast.literal_eval(step_epochs)
        if isinstance(step_epochs, (list, tuple)) and len(step_epochs) > 0:
            epoch_sched = torch.optim.lr_scheduler.MultiStepLR(optimizer=optimizer, milestones=step_epochs, gamma=step_gamma)
    elif schedule_type == 'cosine':
        iter_sched = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer=optimizer, T_max=total_iters, eta_min=0.0)
    elif schedule_type == 'poly':
        iter_sched = PolynomialLR(optimizer=optimizer, T_max=total_iters, power=poly_power, eta_min=0.0)
    else:
        raise ValueError('Unknown schedule_type {}'.format(schedule_type))
    return epoch_sched, iter_sched
