"""Centred padding to a multiple of 2**num_scales and its inverse
(/root/reference/src/helpers/pad.py:23-59).  Host-side shape plumbing."""
from torch.nn import functional as F


def identity(x):
    return x


def padding_tuple(h, w, fac):
    """(left, right, top, bottom); a side whose length is already a multiple gets no padding."""
    need_h, need_w = (-h) % fac, (-w) % fac
    top, left = need_h // 2, need_w // 2
    return (left, need_w - left, top, need_h - top)


def pad(img, fac, mode='replicate'):
    _, _, h, w = img.shape
    t = padding_tuple(h, w, fac)
    if not any(t):
        return img, identity
    assert (t[2] + t[3] + h) % fac == 0 and (t[0] + t[1] + w) % fac == 0
    return F.pad(img, t, mode), t


def undo_pad(img, padLeft, padRight, padTop, padBottom, target_shape=None):
    out = img[..., padTop:(-padBottom or None), padLeft:(-padRight or None)]
    if target_shape:
        assert tuple(out.shape[-2:]) == tuple(target_shape), (out.shape, target_shape)
    return out
