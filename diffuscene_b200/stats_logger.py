"""Running-mean logger with the interface the training loop expects (reference:
scene_synthesis/stats_logger.py:22-125): StatsLogger.instance()[key].value = v accumulates a mean,
print_progress() writes a line, clear() resets.  W&B logging is optional and off when wandb is absent."""
import sys


class AverageAggregator(object):
    def __init__(self):
        self._value = 0.0
        self._count = 0

    @property
    def value(self):
        return self._value / max(self._count, 1)

    @value.setter
    def value(self, val):
        self._value += val
        self._count += 1


class StatsLogger(object):
    _instance = None

    def __init__(self):
        self._values = dict()
        self._loss = AverageAggregator()
        self._output_files = [sys.stdout]

    @classmethod
    def instance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def add_output_file(self, f):
        self._output_files.append(f)

    def __getitem__(self, key):
        if key not in self._values:
            self._values[key] = AverageAggregator()
        return self._values[key]

    def clear(self):
        self._values.clear()
        self._loss = AverageAggregator()

    def print_progress(self, epoch, batch, loss, precision="{:.5f}"):
        self._loss.value = loss
        msg = ("epoch: {} - batch: {} - loss: " + precision).format(epoch, batch, self._loss.value)
        for k, v in self._values.items():
            msg += " - " + k + ": " + precision.format(v.value)
        for f in self._output_files:
            print(msg, flush=True, file=f)
