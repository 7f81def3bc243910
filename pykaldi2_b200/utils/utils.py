"""Meters and LR schedule used by the trainers (reference utils/utils.py:8-53)."""
import time


class AverageMeter(object):
    """Running value / average of a scalar, printed as ``name val (avg)``."""

    def __init__(self, name, fmt=":f"):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = 0.0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / max(self.count, 1)

    def __str__(self):
        return ("{name} {val" + self.fmt + "} ({avg" + self.fmt + "})").format(
            name=self.name, val=self.val, avg=self.avg)


class ProgressMeter(object):
    def __init__(self, num_batches, *meters, prefix=""):
        width = len(str(int(num_batches)))
        self._fmt = "[{:" + str(width) + "d}/" + str(int(num_batches)) + "]"
        self.meters, self.prefix = meters, prefix

    def print(self, batch):
        print("\t".join([self.prefix + self._fmt.format(batch)] + [str(m) for m in self.meters]), flush=True)


def noam_decay(step, warmup_steps, base_lr):
    """Transformer LR schedule (arXiv:1706.03762): base_lr * min(step^-0.5, step * warmup^-1.5)."""
    return base_lr * min(step ** (-0.5), step * warmup_steps ** (-1.5))


class RTFMeter(object):
    """iRTF = hours of audio processed per wall-clock hour (reference README.md:37-39)."""

    def __init__(self):
        self.audio_s, self.t0 = 0.0, time.time()

    def update(self, n_frames, frame_shift_s=0.01):
        self.audio_s += n_frames * frame_shift_s

    @property
    def irtf(self):
        return self.audio_s / max(time.time() - self.t0, 1e-9)
