"""Drop-in for the reference's `torchac` module (/root/reference/src/torchac/torchac.py:87-166):
same four functions, same argument meaning, same exceptions -- backed by libl3c_b200.so instead of
the pybind modules torchac_backend_{cpu,gpu}.

Differences by design: there is no CPU backend (CPU_SUPPORTED is False; `encode_cdf`/`decode_cdf`
still take CPU tensors, as in the reference, but the coding runs on the GPU), and no CDF table is
materialised for `encode_logistic_mixture`.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

CUDA_SUPPORTED = _lib.cuda_supported()
CPU_SUPPORTED = False


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _require_backend():
    if not _lib.cuda_supported():
        raise ValueError('torchac (l3c_pytorch_b200) needs an sm_100 GPU; there is no CPU backend.')


def _check_cdf(cdf):
    if cdf.dim() != 4 or cdf.shape[0] != 1:
        raise RuntimeError('Invalid size for cdf! Expected 1HWLp')        # torchac.cpp:135
    if cdf.dtype != torch.int16:
        raise RuntimeError('cdf must be int16')
    return np.ascontiguousarray(cdf.numpy()).view(np.uint16).reshape(-1, cdf.shape[-1])


def encode_cdf(cdf, sym):
    """cdf: 1HWLp int16 on CPU; sym: int16 on CPU -> bytes."""
    if cdf.is_cuda or sym.is_cuda:
        raise ValueError('CDF and symbols must be on CPU for `encode_cdf`')
    _require_backend()
    table = _check_cdf(cdf)
    s = np.ascontiguousarray(sym.numpy().astype(np.int16, copy=False)).reshape(-1)
    if s.size != table.shape[0]:
        raise RuntimeError('sym has %d entries, cdf describes %d symbols' % (s.size, table.shape[0]))
    cap = s.size * 3 + 64
    out = np.empty(cap, np.uint8)
    n = ctypes.c_size_t(0)
    check(lib.l3c_encode_cdf(_np_ptr(table), table.shape[0], table.shape[1], _np_ptr(s), _np_ptr(out), cap,
                             ctypes.byref(n)))
    return out[:n.value].tobytes()


def decode_cdf(cdf, input_string):
    """cdf: 1HWLp int16 on CPU -> int16 tensor [H*W] on CPU."""
    if cdf.is_cuda:
        raise ValueError('CDF must be on CPU for `decode_cdf`')
    _require_backend()
    table = _check_cdf(cdf)
    buf = np.frombuffer(bytes(input_string), np.uint8)
    out = np.empty(table.shape[0], np.int16)
    check(lib.l3c_decode_cdf(_np_ptr(table), table.shape[0], table.shape[1],
                             _np_ptr(buf) if buf.size else None, buf.size, _np_ptr(out)))
    return torch.from_numpy(out)


def _check_mixture(targets, means, log_scales, logit_probs_softmax):
    if not (targets.is_cuda == means.is_cuda == log_scales.is_cuda == logit_probs_softmax.is_cuda):
        raise ValueError('targets, means, log_scales, logit_probs_softmax must all be on the same device! Got '
                         f'{targets.device}, {means.device}, {log_scales.device}, {logit_probs_softmax.device}.')
    if not targets.is_cuda:
        raise ValueError('Got CPU tensor, but l3c_pytorch_b200 has no CPU backend; move the tensors to the GPU.')
    if means.dim() != 4 or means.shape[0] != 1:
        raise RuntimeError('Invalid size for means! Expected 1KHW')           # torchac.cpp:239
    if means.shape != log_scales.shape or means.shape != logit_probs_softmax.shape:
        raise RuntimeError('Invalid size for log_scales/logit_probs_softmax! Expected 1KHW')
    ts = [t.float().contiguous() for t in (targets, means, log_scales, logit_probs_softmax)]
    return ts, means.shape[1], means.shape[2] * means.shape[3], targets.shape[0]


def encode_logistic_mixture(targets, means, log_scales, logit_probs_softmax, sym):
    """targets [Lp], means/log_scales/logit_probs_softmax 1KHW (CUDA); sym int16 on CPU -> bytes."""
    (t, m, ls, pi), K, n_sym, Lp = _check_mixture(targets, means, log_scales, logit_probs_softmax)
    if sym.is_cuda:
        raise ValueError('sym must be on CPU!')
    s = np.ascontiguousarray(sym.numpy().astype(np.int16, copy=False)).reshape(-1)
    if s.size != n_sym:
        raise RuntimeError('sym has %d entries, parameters describe %d symbols' % (s.size, n_sym))
    cap = n_sym * 3 + 64
    out = np.empty(cap, np.uint8)
    n = ctypes.c_size_t(0)
    torch.cuda.current_stream().synchronize()       # inputs may have been produced on another stream
    check(lib.l3c_encode_logistic_mixture(t.data_ptr(), m.data_ptr(), ls.data_ptr(), pi.data_ptr(), K, n_sym, Lp,
                                          _np_ptr(s), _np_ptr(out), cap, ctypes.byref(n)))
    return out[:n.value].tobytes()


def decode_logistic_mixture(targets, means, log_scales, logit_probs_softmax, input_string):
    (t, m, ls, pi), K, n_sym, Lp = _check_mixture(targets, means, log_scales, logit_probs_softmax)
    buf = np.frombuffer(bytes(input_string), np.uint8)
    out = np.empty(n_sym, np.int16)
    torch.cuda.current_stream().synchronize()
    check(lib.l3c_decode_logistic_mixture(t.data_ptr(), m.data_ptr(), ls.data_ptr(), pi.data_ptr(), K, n_sym, Lp,
                                          _np_ptr(buf) if buf.size else None, buf.size, _np_ptr(out)))
    return torch.from_numpy(out)
