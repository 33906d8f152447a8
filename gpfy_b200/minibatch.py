"""Shuffled minibatches on the device (mirror of `batched_data_generator`, optimization.py:31-58).

The reference wraps a HuggingFace Dataset: `dataset.shuffle().iter(batch_size, drop_last_batch=True)`, restarted when
exhausted.  Here the rows stay where they are - in HBM, or in pinned host memory when the data set exceeds it - and a
batch is gathered by ONE kernel (`gpfy_minibatch_gather`) that evaluates the epoch's permutation on the fly (a keyed
Feistel bijection of [0, N); no permutation array, no sort, no host round trip).  Every batch of an epoch is disjoint
from the others, the last partial batch is dropped, and a new epoch draws a new permutation, as in the reference.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import _as_dev, _ptr, _stream, require_cuda


def shuffle_index(seed, epoch, n_rows, position):
    """pi_{seed, epoch}(position): the data row at `position` of the epoch (host evaluation of the kernel's function)."""
    out = ctypes.c_int64(0)
    _lib.check(_lib.load().gpfy_shuffle_index(int(seed), int(epoch), int(n_rows), int(position), ctypes.byref(out)))
    return out.value


def _resident(x, device):
    """Device tensors and pinned host tensors are used in place; anything else is copied ONCE, here: to the device when it
    fits in 40 % of the free memory, otherwise into pinned host memory (the gather kernel reads it through unified addressing)."""
    if torch.is_tensor(x) and x.dtype == torch.float64 and x.is_contiguous() and (x.is_cuda or x.is_pinned()):
        return x
    t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64)))
    if not t.is_cuda and t.numel() * 8 > 0.4 * torch.cuda.mem_get_info(device)[0]:
        return t.to(torch.float64).contiguous().pin_memory()
    return _as_dev(t, device)


class DeviceMinibatches:
    """Iterator over (X_batch, Y_batch) device tensors.  The two output slabs are reused: a batch is valid until the
    next `next()` on the same stream (which is how a training step consumes it)."""

    def __init__(self, X, Y, batch_size, seed=0, device=None, return_index=False):
        self.lib = _lib.load()
        self.device = require_cuda(device)
        self.X, self.Y = _resident(X, self.device), _resident(Y, self.device)
        if self.Y.dim() == 1:
            self.Y = self.Y[:, None]
        self.n = int(self.X.shape[0])
        self.dx, self.dy = int(self.X.shape[1]), int(self.Y.shape[1])
        self.batch_size = int(batch_size)
        if not 0 < self.batch_size <= self.n:
            raise _lib.GpfyError(f"batch_size {batch_size} outside (0, {self.n}]")
        if int(self.Y.shape[0]) != self.n:
            raise _lib.GpfyError("X and Y disagree on the number of rows")
        self.seed, self.epoch, self.pos = int(seed), 0, 0
        self.Xb = torch.empty((self.batch_size, self.dx), dtype=torch.float64, device=self.device)
        self.Yb = torch.empty((self.batch_size, self.dy), dtype=torch.float64, device=self.device)
        self.idx = torch.empty(self.batch_size, dtype=torch.int64, device=self.device) if return_index else None

    def __iter__(self):
        return self

    def __next__(self):
        if self.pos + self.batch_size > self.n:      # drop_last_batch=True, then reshuffle (optimization.py:48-57)
            self.epoch, self.pos = self.epoch + 1, 0
        _lib.check(self.lib.gpfy_minibatch_gather(_stream(self.device), self.seed, self.epoch, self.n, self.pos, self.batch_size,
                                                  _ptr(self.X), self.dx, _ptr(self.Y), self.dy, _ptr(self.Xb), _ptr(self.Yb),
                                                  _ptr(self.idx)))
        self.pos += self.batch_size
        return self.Xb, self.Yb

    next = __next__

    def next_from_state(self, state):
        """The batch at (epoch, position) = state[1:3] on the device (gpfy_step_advance moves it on): no host-side
        bookkeeping, so the call can sit inside a CUDA graph."""
        _lib.check(self.lib.gpfy_minibatch_gather_state(_stream(self.device), self.seed, self.n, self.batch_size, _ptr(state),
                                                        _ptr(self.X), self.dx, _ptr(self.Y), self.dy, _ptr(self.Xb), _ptr(self.Yb),
                                                        _ptr(self.idx)))
        return self.Xb, self.Yb
