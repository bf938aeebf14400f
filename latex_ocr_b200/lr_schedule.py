"""LRSchedule — host-side learning-rate schedule with the semantics of model/utils/lr_schedule.py:27-118 (warm start,
exponential decay between start_decay and end_decay, multiplicative decay / early stopping on a non-improving score,
floor at lr_min; all durations in batches).  Pure host scalar logic, restated as a small state machine.

Note: the reference's torch trainer only *displays* this value (img2seq_torch.py:121); pass
``apply_to=model`` to actually drive the fused Adam's learning rate (device scalar, no graph re-capture needed)."""
import math


class LRSchedule:
    def __init__(self, lr_init=1e-3, lr_min=1e-4, start_decay=0, decay_rate=None, end_decay=None, lr_warm=1e-4,
                 end_warm=None, early_stopping=None, apply_to=None):
        self._lr_init, self._lr_min = lr_init, lr_min
        self._decay_rate, self._end_decay = decay_rate, end_decay
        self._lr_warm, self._end_warm = lr_warm, end_warm
        self._early_stopping = early_stopping
        self._score = None
        self._n_batch_no_imprv = 0
        self._apply_to = apply_to
        warm = end_warm is not None
        self._start_decay = max(end_warm, start_decay) if warm else start_decay       # decay never starts inside the warm-up
        self.lr = lr_warm if warm else lr_init
        if end_decay is not None:
            self._exp_decay = math.pow(lr_min / lr_init, 1.0 / float(end_decay - self._start_decay))

    @property
    def stop_training(self):
        return self._early_stopping is not None and self._n_batch_no_imprv >= self._early_stopping

    def update(self, batch_no=None, score=None):
        if batch_no is not None:
            if self._end_warm is not None and self._end_warm <= batch_no <= self._start_decay:
                self.lr = self._lr_init
            if self._end_decay is not None and batch_no > self._start_decay:
                self.lr *= self._exp_decay
        if self._decay_rate is not None and score is not None and self._score is not None:
            if score <= self._score:
                self.lr *= self._decay_rate
                self._n_batch_no_imprv += 1
            else:
                self._n_batch_no_imprv = 0
        if score is not None:
            self._score = score
        self.lr = max(self.lr, self._lr_min)
        if self._apply_to is not None:
            self._apply_to.set_lr(self.lr)
