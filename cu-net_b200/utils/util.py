"""Drop-in for the pieces of the reference's ``utils/util.py`` that the train / validate loop touches:
``TrainHistory`` (:8-46), ``AverageMeter`` (:86-104) and the ``adjust_lr`` schedule (:106-119).  Host bookkeeping
only -- no arithmetic of the hot path lives here."""
from collections import OrderedDict


class TrainHistory(object):
    """utils/util.py:8-46, key for key: the per-epoch records are lists of (Ordered)dicts
    ``{'epoch': e}``, ``{'lr': lr}``, ``{'train_loss':, 'val_loss':}``, ``{'val_pckh':}`` plus ``best_pckh`` / ``is_best``
    -- exactly what the reference's ``load_state_dict`` (:40-46) reads, so a checkpoint written here resumes there and
    vice versa (tests/test_checkpoint_cpu.py round-trips it against the real class)."""

    def __init__(self):
        self.epoch, self.lr, self.loss, self.pckh = [], [], [], []
        self.best_pckh = 0.
        self.is_best = True

    def update(self, epoch, lr, loss, pckh):
        self.epoch.append(epoch)
        self.lr.append(lr)
        self.loss.append(loss)
        self.pckh.append(pckh)
        self.is_best = pckh['val_pckh'] > self.best_pckh
        self.best_pckh = max(pckh['val_pckh'], self.best_pckh)

    def state_dict(self):
        return OrderedDict([('epoch', self.epoch), ('lr', self.lr), ('loss', self.loss), ('pckh', self.pckh),
                            ('best_pckh', self.best_pckh), ('is_best', self.is_best)])

    def load_state_dict(self, state_dict):
        self.epoch, self.lr = list(state_dict['epoch']), list(state_dict['lr'])
        # files written by round-1 builds of this repo carried only epoch / lr: fill what is missing
        self.loss = list(state_dict.get('loss', [OrderedDict([('train_loss', 0.), ('val_loss', 0.)])] * len(self.epoch)))
        self.pckh = list(state_dict.get('pckh', [OrderedDict([('val_pckh', 0.)])] * len(self.epoch)))
        self.best_pckh = state_dict.get('best_pckh', 0.)
        self.is_best = state_dict.get('is_best', True)

    def last_epoch(self):
        """Epoch index of the newest record, -1 for an empty history (start_epoch = last + 1, cu-net.py:112-114)."""
        return self.epoch[-1]['epoch'] if self.epoch else -1


class AverageMeter(object):
    """utils/util.py:86-104."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def adjust_lr(opt, epoch, optimizer=None):
    """utils/util.py:106-119: x0.2 at epoch 101, x0.5 at 141 and 161; also writes the optimizer's param groups."""
    if epoch == 101:
        opt.lr *= 0.2
    elif epoch in (141, 161):
        opt.lr *= 0.5
    if optimizer is not None:
        for group in optimizer.param_groups:
            group["lr"] = opt.lr
    return opt.lr
