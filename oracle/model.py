"""oracle/model.py -- TEST INFRASTRUCTURE ONLY.

Functional CPU restatement (plain PyTorch fp32 ops + the C oracle of ac_oracle.c) of the
reference's encode/decode hot path, driven by a reference-layout state dict.  It is the checker
for the CUDA path; the product never imports it.

Reference anchors (all under /root/reference/src):
  network forward / get_P ........ modules/multiscale_network.py:226-322, modules/net.py:89-184,
                                   modules/edsr.py:52-119, modules/head.py:26-59,
                                   modules/prob_clf.py:29-74, pytorch_ext.py:57-61
  quantizer (eval path) .......... modules/quantizer.py:38-90
  DMLL NLL ....................... criterion/logistic_mixture.py:146-246,334-345
  per-channel CDF parameters ..... criterion/logistic_mixture.py:134-141,248-275
  CDF table ...................... torchac/torchac.py:174-213 (PyTorch path) and
                                   torchac/torchac_backend/torchac_kernel.cu:20-76 (kernel formula)
  bit-coding loop + container .... bitcoding/bitcoding.py:50-375, bitcoding/coders.py:33-90
  pad / crops .................... helpers/pad.py:23-59, auto_crop.py:44-152

Two decode-side repairs are built in (SURVEY.md findings 1 and 2; without them the unmodified
reference cannot round-trip): the bit-decoder maps symbols to values through the same `levels`
LUT the encoder's quantizer uses, and the RGB baselines subtract rgb_mean from the decoded
thumbnail.  Encoder-side bytes are unaffected by either.
"""
import io
import struct
from collections import namedtuple

import numpy as np
import torch
import torch.nn.functional as F

from . import ac

MAGIC = b'\x46\xE2\x84\x92'          # bitcoding.py:36
LOG_SCALES_MIN = -7.0                # logistic_mixture.py:57

Cfg = namedtuple('Cfg', ['num_scales', 'Cf', 'K', 'qC', 'qL', 'levels_range', 'n_blocks_enc',
                         'n_blocks_dec', 'rgb_baseline', 'dec_skip', 'feed_F'])

CFG_L3C = Cfg(3, 64, 10, 5, 25, (-1.0, 1.0), 8, 8, False, True, True)            # configs/ms/cr.cf
CFG_RGB_SHARED = Cfg(1, 64, 10, 3, 5, (-1.0, 1.0), 8, 8, True, False, False)     # cr_rgb_shared.cf
RGB_MEAN = (0.4488, 0.4371, 0.4040)


# ----------------------------------------------------------------------------------------------
# network
# ----------------------------------------------------------------------------------------------
def _conv(sd, name, x, stride=1, rate=1):
    w = sd[name + '.weight']
    k = w.shape[-1]
    pad = k // 2 if rate == 1 else rate                      # pytorch_ext.py:57-61
    return F.conv2d(x, w, sd[name + '.bias'], stride=stride, padding=pad, dilation=rate)


def _body(sd, prefix, x, n_blocks):
    """n ResBlocks (conv-ReLU-conv, += x) + conv, with the outer skip (net.py:144, 181)."""
    y = x
    for b in range(n_blocks):
        r = _conv(sd, '%s.%d.body.0' % (prefix, b), y)
        r = F.relu(r)
        r = _conv(sd, '%s.%d.body.2' % (prefix, b), r)
        y = r + y
    y = _conv(sd, '%s.%d' % (prefix, n_blocks), y)
    return y + x


def quantize(q_in, levels):
    """hard path of Quantizer.forward (quantizer.py:62-90): argmin_l (x-level_l)^2, levels[S]."""
    d = (q_in.unsqueeze(-1) - levels) ** 2
    S = torch.min(d, dim=-1)[1]
    return S, levels[S]


def _rgb_mean255():
    return torch.tensor(RGB_MEAN, dtype=torch.float32).reshape(3, 1, 1).mul(255.)


def pillow_bicubic_half(u8_nchw):
    """resize_bicubic_batch(t, 0.5) (dataloaders/images_loader.py:277-293): Pillow BICUBIC on
    uint8 HWC.  NOTE the reference passes (int(h*fac), int(w*fac)) with h,w = img.size = (W,H),
    i.e. the usual (W/2, H/2)."""
    from PIL import Image
    outs = []
    for t in u8_nchw:
        img = Image.fromarray(t.permute(1, 2, 0).contiguous().numpy())
        w, h = img.size
        img = img.resize((int(w * 0.5), int(h * 0.5)), Image.BICUBIC)
        outs.append(torch.from_numpy(np.array(img)).permute(2, 0, 1))
    return torch.stack(outs, 0)


def encoder_side(sd, cfg, x_sub):
    """fine->coarse loop of _forward_with_scales (multiscale_network.py:276-283).
    Returns per-scale lists S (symbols, int64), bn_q (float)."""
    S, bnq = [], []
    inp = x_sub
    for s in range(cfg.num_scales):
        if cfg.rgb_baseline:
            m = _rgb_mean255()
            u8 = (inp + m).clamp(0, 255.).round().to(torch.uint8)       # net.py:72-75
            down = pillow_bicubic_half(u8)
            S.append(down.long())
            bnq.append(down.float() - m)
            inp = bnq[-1]
            continue
        if s == 0:
            h = _conv(sd, 'heads.0.head.0', inp)                          # MeanShift 1/128
            h = _conv(sd, 'heads.0.head.1.head', h)
        else:
            h = _conv(sd, 'heads.%d.head' % s, inp)
        p = 'nets.%d.enc' % s
        x = _conv(sd, p + '.down', h, stride=2)
        x = _body(sd, p + '.body', x, cfg.n_blocks_enc)
        q_in = _conv(sd, p + '.to_q.0', x)
        sym, hard = quantize(q_in, sd[p + '.levels'])
        S.append(sym)
        bnq.append(hard)
        inp = x                                                           # feed_F
    return S, bnq


def decoder_net(sd, cfg, scale, bn_q, F_prev):
    """EDSRDec.forward (net.py:173-184) -> F at the resolution of `scale`."""
    p = 'nets.%d.dec' % scale
    x = _conv(sd, p + '.head', bn_q)
    if F_prev is not None:
        x = x + F_prev
    x = _body(sd, p + '.body', x, cfg.n_blocks_dec)
    x = _conv(sd, p + '.tail.0', x)
    return F.pixel_shuffle(x, 2)


def prob_clf(sd, scale, Fdec):
    """StackedAtrousConvs.forward (prob_clf.py:70-73)."""
    p = 'prob_clfs.%d.atrous' % scale
    cat = torch.cat([_conv(sd, '%s.atrous.%d' % (p, i), Fdec, rate=r)
                     for i, r in enumerate((1, 2, 4))], dim=1)
    return _conv(sd, p + '.lin', cat)


def get_P(sd, cfg, scale, bn_q, F_prev):
    """MultiscaleNetwork.get_P (multiscale_network.py:308-322)."""
    Fd = decoder_net(sd, cfg, scale, bn_q, F_prev if cfg.dec_skip else None)
    return prob_clf(sd, scale, Fd), Fd


Out = namedtuple('Out', ['S', 'bn', 'P', 'L'])


def forward(sd, cfg, img):
    """MultiscaleNetwork.forward in eval mode (multiscale_network.py:226-306).
    img: N x 3 x H x W float (0..255). Returns Out with S/bn fine->coarse (len scales+1), P (len scales)."""
    with torch.no_grad():
        S0 = img.round().long()
        x = _conv(sd, 'sub_rgb_mean', img)
        S, bnq = encoder_side(sd, cfg, x)
        P = [None] * cfg.num_scales
        F_prev = None
        for s in reversed(range(cfg.num_scales)):
            P[s], F_prev = get_P(sd, cfg, s, bnq[s], F_prev)
        L_other = 256 if cfg.rgb_baseline else cfg.qL
        return Out([S0] + S, [None] + bnq, P, [256] + [L_other] * cfg.num_scales)


# ----------------------------------------------------------------------------------------------
# DMLL
# ----------------------------------------------------------------------------------------------
class Dmll(object):
    """Value grid of DiscretizedMixLogisticLoss (logistic_mixture.py:86-124)."""
    def __init__(self, rgb, x_min, x_max, L):
        self.rgb, self.x_min, self.x_max, self.L = rgb, float(x_min), float(x_max), L
        self.bin_width = (x_max - x_min) / (L - 1)
        self.num_params = 4 if rgb else 3

    def targets(self):
        return torch.linspace(self.x_min - self.bin_width / 2, self.x_max + self.bin_width / 2,
                              self.L + 1, dtype=torch.float32)      # coders_helpers.py:44-46


def dmlls(cfg):
    rgb = Dmll(True, 0, 255, 256)
    return rgb, (rgb if cfg.rgb_baseline else Dmll(False, cfg.levels_range[0], cfg.levels_range[1], cfg.qL))


def nll(dm, x, l):
    """DiscretizedMixLogisticLoss.forward (logistic_mixture.py:146-246): NCHW nats."""
    N, C, H, W = x.shape
    K = l.shape[1] // (dm.num_params * C)
    l = l.reshape(N, dm.num_params, C, K, H, W)
    logit_pis, means = l[:, 0], l[:, 1]
    log_scales = torch.clamp(l[:, 2], min=LOG_SCALES_MIN)
    x = x.reshape(N, C, 1, H, W)
    if dm.rgb:
        co = torch.sigmoid(l[:, 3])
        means = torch.stack((means[:, 0],
                             means[:, 1] + co[:, 0] * x[:, 0],
                             means[:, 2] + co[:, 1] * x[:, 0] + co[:, 2] * x[:, 1]), dim=1)
    cx = x - means
    inv = torch.exp(-log_scales)
    plus_in = inv * (cx + dm.bin_width / 2)
    min_in = inv * (cx - dm.bin_width / 2)
    cdf_plus, cdf_min = torch.sigmoid(plus_in), torch.sigmoid(min_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)
    log_one_minus_cdf_min = -F.softplus(min_in)
    out_A = torch.log(torch.clamp(cdf_plus - cdf_min, min=1e-12))
    cond_B = (x > dm.x_max - 0.001).float()
    out_B = cond_B * log_one_minus_cdf_min + (1. - cond_B) * out_A
    cond_C = (x < dm.x_min + 0.001).float()
    log_probs = cond_C * log_cdf_plus + (1. - cond_C) * out_B
    lw = log_probs + torch.log_softmax(logit_pis, dim=2)
    return -torch.logsumexp(lw, dim=2)


def theoretical_bpsps(cfg, out, num_subpixels=None):
    """Losses.get + MultiscaleBlueprint.get_loss (multiscale_network.py:145-165,
    multiscale_blueprint.py:64-95): per-scale bpsp incl. the uniform-prior final scale."""
    rgb, other = dmlls(cfg)
    costs = [float(nll(rgb, out.S[0].float(), out.P[0]).sum())]
    for s in range(1, cfg.num_scales):
        tgt = out.S[s].float() if cfg.rgb_baseline else out.bn[s]
        costs.append(float(nll(other, tgt, out.P[s]).sum()))
    n_sub = num_subpixels or int(np.prod(out.S[0].shape))
    conv = np.log(2.) * n_sub
    final = float(np.prod(out.S[-1].shape)) * np.log(out.L[-1])
    return [c / conv for c in costs] + [final / conv]


def channel_params(dm, l, c, C, x_dec):
    """cdf_step_non_shared / _extract_non_shared_c (logistic_mixture.py:134-141,248-275).
    l: 1 x Kp x H x W.  x_dec: 1 x C x H x W values of the channels coded so far.
    Returns (softmax pi, mu, clamped log sigma), each K x (H*W) float32."""
    N, Kp, H, W = l.shape
    K = Kp // (dm.num_params * C)
    l = l.reshape(N, dm.num_params, C, K, H, W)
    pi = F.softmax(l[:, 0, c], dim=1)
    mu = l[:, 1, c].clone()
    ls = torch.clamp(l[:, 2, c], min=LOG_SCALES_MIN)
    if dm.rgb and c == 1:
        mu = mu + torch.sigmoid(l[:, 3, 0]) * x_dec[:, 0]
    elif dm.rgb and c == 2:
        mu = mu + (torch.sigmoid(l[:, 3, 1]) * x_dec[:, 0] + torch.sigmoid(l[:, 3, 2]) * x_dec[:, 1])
    return [t.reshape(K, H * W).contiguous().numpy() for t in (pi, mu, ls)]


def cdf_table_kernel_formula(dm, l, c, C, x_dec):
    """uint16 [H*W, L+1] table by the reference CUDA kernel's formula (torchac_kernel.cu:20-76)."""
    pi, mu, ls = channel_params(dm, l, c, C, x_dec)
    return ac.mixture_cdf(dm.targets().numpy(), mu, ls, pi)


def cdf_table_torch_formula(dm, l, c, C, x_dec):
    """uint16 [H*W, L+1] table by the reference's PyTorch CPU path (torchac.py:174-213)."""
    pi, mu, ls = [torch.from_numpy(a) for a in channel_params(dm, l, c, C, x_dec)]
    t = dm.targets()
    cdf = ((t - mu.unsqueeze(-1)) * torch.exp(-ls).unsqueeze(-1)).sigmoid()
    cdf = (cdf * pi.unsqueeze(-1)).sum(0)
    Lp = t.numel()
    cdf = (cdf * (65536.0 - (Lp - 1))).round().to(torch.int16) + torch.arange(Lp, dtype=torch.int16)
    return cdf.numpy().view(np.uint16)


# ----------------------------------------------------------------------------------------------
# pad / crops
# ----------------------------------------------------------------------------------------------
def pad_tuple(h, w, fac):
    """helpers/pad.py:23-46 -> (left, right, top, bottom)."""
    ph, pw = fac - (h % fac), fac - (w % fac)
    if ph == fac and pw == fac:
        return (0, 0, 0, 0)
    t = 0 if ph == fac else ph // 2
    b = 0 if ph == fac else ph - t
    le = 0 if pw == fac else pw // 2
    r = 0 if pw == fac else pw - le
    return (le, r, t, b)


def iter_crops(img, needs_crop_dim):
    """auto_crop.py:55-88 (depth-first TL,TR,BL,BR)."""
    H, W = img.shape[-2:]
    if H * W <= needs_crop_dim:
        yield img
        return
    for part in (img[..., :H // 2, :W // 2], img[..., :H // 2, W // 2:],
                 img[..., H // 2:, :W // 2], img[..., H // 2:, W // 2:]):
        yield from iter_crops(part, needs_crop_dim)


# ----------------------------------------------------------------------------------------------
# bit-coding (one image, one file)
# ----------------------------------------------------------------------------------------------
def encode_image(sd, cfg, img_u8, cdf_formula='kernel', return_debug=False):
    """Bitcoding.encode for an image that needs no crop (bitcoding.py:73-123).
    img_u8: 3 x H x W uint8 tensor.  Returns container bytes (+ debug dict)."""
    build = cdf_table_kernel_formula if cdf_formula == 'kernel' else cdf_table_torch_formula
    rgb, other = dmlls(cfg)
    img = img_u8.unsqueeze(0).long()
    fac = 2 ** cfg.num_scales
    _, _, H, W = img.shape
    pt = pad_tuple(H, W, fac)
    if any(pt):
        img = F.pad(img, pt, 'constant')
    imgf = img.float()
    out = forward(sd, cfg, imgf)
    f = io.BytesIO()
    f.write(struct.pack('<4H', *pt))
    streams = {}
    for scale in reversed(range(cfg.num_scales + 1)):
        S = out.S[scale]
        _, C, h, w = S.shape
        f.write(struct.pack('<BHH', C, h, w))
        dm = rgb if scale == 0 else other
        if scale == cfg.num_scales:
            row = ac.uniform_cdf_row(dm.L)
            for c in range(C):
                b = ac.encode(row, S[0, c].reshape(-1).numpy().astype(np.int16))
                streams[(scale, c)] = b
                f.write(struct.pack('<I', len(b)))
                f.write(b)
        else:
            l = out.P[scale]
            bn = imgf if scale == 0 else out.bn[scale]
            dec = torch.zeros_like(bn)
            for c in range(C):
                table = build(dm, l, c, C, dec)
                b = ac.encode(table, S[0, c].reshape(-1).numpy().astype(np.int16))
                streams[(scale, c)] = b
                f.write(struct.pack('<I', len(b)))
                f.write(b)
                dec[:, c] = bn[:, c]
        f.write(MAGIC)
    data = f.getvalue()
    if return_debug:
        return data, dict(out=out, streams=streams, pad=pt,
                          bpsp=len(data) * 8 / float(np.prod(img.shape)))
    return data


def decode_image(sd, cfg, data, cdf_formula='kernel'):
    """Bitcoding.decode for a single (non-part) file (bitcoding.py:125-161), with the two
    decode-side repairs described in the module docstring.  Returns 1 x 3 x H x W int64."""
    build = cdf_table_kernel_formula if cdf_formula == 'kernel' else cdf_table_torch_formula
    rgb, other = dmlls(cfg)
    f = io.BytesIO(data)
    pt = struct.unpack('<4H', f.read(8))
    bn_prev, F_prev = None, None
    with torch.no_grad():
        for scale in reversed(range(cfg.num_scales + 1)):
            C, h, w = struct.unpack('<BHH', f.read(5))
            dm = rgb if scale == 0 else other
            if scale == cfg.num_scales:
                row = ac.uniform_cdf_row(dm.L)
                S = []
                for c in range(C):
                    n, = struct.unpack('<I', f.read(4))
                    S.append(torch.from_numpy(ac.decode(row, f.read(n), h * w).astype(np.int64)))
                S = torch.stack(S, 0).reshape(1, C, h, w)
                if cfg.rgb_baseline:
                    bn_prev = S.float() - _rgb_mean255()
                else:
                    bn_prev = torch.linspace(cfg.levels_range[0], cfg.levels_range[1], cfg.qL)[S]
            else:
                l, F_prev = get_P(sd, cfg, scale, bn_prev, F_prev)
                dec = torch.zeros(1, C, h, w)
                for c in range(C):
                    table = build(dm, l, c, C, dec)
                    n, = struct.unpack('<I', f.read(4))
                    s_c = torch.from_numpy(ac.decode(table, f.read(n)).astype(np.int64)).reshape(h, w)
                    if scale == 0 or cfg.rgb_baseline:
                        dec[0, c] = s_c.float()
                    else:
                        dec[0, c] = torch.linspace(cfg.levels_range[0], cfg.levels_range[1], cfg.qL)[s_c]
                bn_prev = dec
            assert f.read(4) == MAGIC
    out = bn_prev.round().long()
    le, r, t, b = pt
    return out[..., t:(-b or None), le:(-r or None)]
