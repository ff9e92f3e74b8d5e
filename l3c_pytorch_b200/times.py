"""Timing scopes with the duck type `Bitcoding` expects from the reference's StackTimeLogger
(/root/reference/src/test/cuda_timer.py:107-151): `run(name)`, `prefix_scope(p)`, `combine(fmt)`,
`skip(flag)` context managers; device work is synchronised around every scope unless
NO_CUDA_SYNC=1."""
import os
import time
from collections import OrderedDict
from contextlib import contextmanager

import torch


class _NoOpTimes(object):
    """Stand-in for fjcommon.no_op.NoOp: everything is a no-op context manager."""

    def __getattr__(self, _):
        return self

    def __call__(self, *a, **kw):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


NoOp = _NoOpTimes()


def _sync():
    if os.environ.get('NO_CUDA_SYNC', '0') != '1' and torch.cuda.is_available():
        torch.cuda.synchronize()


class StackTimeLogger(object):
    def __init__(self):
        self.records = OrderedDict()     # name -> [seconds]
        self._prefix = []
        self._skip = False

    @contextmanager
    def skip(self, flag):
        old, self._skip = self._skip, bool(flag)
        try:
            yield
        finally:
            self._skip = old

    @contextmanager
    def prefix_scope(self, p):
        self._prefix.append(p)
        try:
            yield
        finally:
            self._prefix.pop()

    @contextmanager
    def combine(self, fmt=None):
        yield

    @contextmanager
    def run(self, name):
        _sync()
        t0 = time.perf_counter()
        try:
            yield
        finally:
            _sync()
            if not self._skip:
                key = ' '.join(self._prefix + [name])
                self.records.setdefault(key, []).append(time.perf_counter() - t0)

    def get_mean_strs(self):
        return ['{}: {:.5f}'.format(k, sum(v) / len(v)) for k, v in self.records.items()]

    def get_last_strs(self):
        return ['{}: {:.5f}'.format(k, v[-1]) for k, v in self.records.items()]
