"""Training step of L3C -- the one part of this package that is NOT hand-written
CUDA: a plain PyTorch-autograd restatement over the SAME parameter tree (`MultiscaleNetwork`'s nn.Conv2d
containers), so that what it trains loads into the sm_100a coding path unchanged.  It exists so that a user of the
reference's `train.py` finds the step here; it is not on the encode/decode hot path and is not benchmarked.

Reference (under /root/reference/src):
  train/multiscale_trainer.py:173-226   train_step: zero_grad, forward, get_loss, backward, optim.step
  modules/quantizer.py:62-90            soft/hard quantiser (straight-through: hard values, soft gradient)
  modules/net.py:136-148,173-184        encoder / decoder forward (training: the decoder is fed `bn`, the soft one)
  modules/multiscale_network.py:262-306 _forward_with_scales
  criterion/logistic_mixture.py:146-246 discretised-logistic-mixture NLL (+ RGB mean coupling)
  blueprints/multiscale_blueprint.py:64-95  bpsp conversion, loss_pc = sum of the per-scale costs without the
                                            uniform-prior final scale
Works on CPU tensors as well as CUDA ones (it only uses torch operators).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .dmll import non_shared_get_K

_LOG_SCALES_MIN = -7.


def soft_quantize(x, levels, sigma):
    """quantizer.py:62-90 -> (x_soft with hard data and soft gradient, x_hard, symbols)."""
    N, C, H, W = x.shape
    xv = x.view(N, C, H * W, 1)
    d = torch.pow(xv - levels, 2)
    phi_soft = F.softmax(-sigma * d, dim=-1)
    x_soft = torch.sum(levels * phi_soft, dim=-1).view(N, C, H, W)
    _, sym = torch.min(d.detach(), dim=-1)
    sym = sym.view(N, C, H, W)
    x_hard = levels[sym]
    x_soft.data = x_hard                      # assign data, keep gradient
    return x_soft, x_hard, sym


def _res_body(body, x):
    """Sequential of ResBlocks + final conv with the outer skip (net.py:144,181; edsr.py:83-86)."""
    y = x
    for m in list(body)[:-1]:
        y = m.body[2](F.relu(m.body[0](y))) + y
    return body[-1](y) + x


def _enc_forward(enc, x):
    """EDSRLikeEnc.forward in training mode (net.py:136-148): -> (bn soft, bn_q, S, F)."""
    x = enc.down(x)
    x = _res_body(enc.body, x)
    feat = x
    q = enc.to_q[0](x)
    bn, bn_q, sym = soft_quantize(q, enc.levels, enc.q.sigma)
    return bn, bn_q, sym, feat


def _dec_forward(dec, bn, fuse):
    x = dec.head(bn)
    if fuse is not None:
        x = x + fuse
    x = _res_body(dec.body, x)
    return F.pixel_shuffle(dec.tail[0](x), 2)


def _prob_clf(clf, x):
    a = clf.atrous
    return a.lin(torch.cat([c(x) for c in a.atrous], dim=1))


def dmll_nll(dm, x, l):
    """DiscretizedMixLogisticLoss.forward (logistic_mixture.py:146-207): x NCHW targets on the value grid (may carry
    gradient), l NKpHW -> NCHW negative log-likelihood in nats."""
    N, C, H, W = x.shape
    K = non_shared_get_K(l.shape[1], C)
    lr = l.reshape(N, dm._num_params, C, K, H, W)
    logit_pis, means = lr[:, 0], lr[:, 1]
    log_scales = torch.clamp(lr[:, 2], min=_LOG_SCALES_MIN)
    xr = x.reshape(N, C, 1, H, W)
    if dm.use_coeffs:
        assert C == 3
        co = torch.sigmoid(lr[:, 3])
        means = torch.stack((means[:, 0],
                             means[:, 1] + co[:, 0] * xr[:, 0],
                             means[:, 2] + co[:, 1] * xr[:, 0] + co[:, 2] * xr[:, 1]), dim=1)
    centered = xr - means
    inv_stdv = torch.exp(-log_scales)
    plus_in = inv_stdv * (centered + dm.bin_width / 2)
    min_in = inv_stdv * (centered - dm.bin_width / 2)
    cdf_delta = torch.sigmoid(plus_in) - torch.sigmoid(min_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)
    log_one_minus_cdf_min = -F.softplus(min_in)
    out_a = torch.log(torch.clamp(cdf_delta, min=1e-12))
    cond_b = (xr > dm.x_upper_bound).float()
    out_b = cond_b * log_one_minus_cdf_min + (1. - cond_b) * out_a
    cond_c = (xr < dm.x_lower_bound).float()
    log_probs = cond_c * log_cdf_plus + (1. - cond_c) * out_b
    weighted = log_probs + F.log_softmax(logit_pis, dim=2)
    return -torch.logsumexp(weighted, dim=2)


def forward_train(net, x):
    """MultiscaleNetwork.forward in training mode (multiscale_network.py:226-306) for the L3C configurations:
    x NCHW float in [0, 255] -> (targets, predictions) per scale fine -> coarse, plus the symbols of every scale.
    targets[0] is the image, targets[i >= 1] the SOFT bottleneck of scale i (gradient flows through it)."""
    if net._rgb:
        raise NotImplementedError('the training step is built for the L3C configurations (learned bottlenecks)')
    syms = [x.round().long()]
    h = net.sub_rgb_mean(x)
    inp = h
    encs = []
    for s in range(net.scales):
        head = net.heads[s]
        inp = head.head[1].head(head.head[0](inp)) if s == 0 else head.head(inp)
        bn, bn_q, sym, feat = _enc_forward(net.nets[s].enc, inp)
        encs.append((bn, bn_q))
        syms.append(sym)
        inp = feat                                   # enc.feed_F
    dec_f = [None] * net.scales
    prev = None
    for s in reversed(range(net.scales)):
        fuse = prev if (net._fuse_feat and s != net.scales - 1) else None
        prev = _dec_forward(net.nets[s].dec, encs[s][0], fuse)          # training: the decoder sees the soft bn
        dec_f[s] = prev
    preds = [_prob_clf(net.prob_clfs[s], dec_f[s]) for s in range(net.scales)]
    targets = [x] + [encs[s][0] for s in range(net.scales - 1)]
    return targets, preds, syms


def training_loss(blueprint, x):
    """-> (loss_pc, [bpsp per scale incl. the uniform-prior final scale]) as multiscale_blueprint.py:64-95."""
    net, losses = blueprint.net, blueprint.losses
    targets, preds, syms = forward_train(net, x)
    costs = []
    for i, (t, p) in enumerate(zip(targets, preds)):
        dm = losses.loss_dmol_rgb if i == 0 else losses.loss_dmol_n
        costs.append(dmll_nll(dm, t, p).sum())
    num_subpixels = int(np.prod(x.shape))
    conversion = np.log(2.) * num_subpixels
    costs_bpsp = [c / conversion for c in costs]
    final = int(np.prod(syms[-1].shape)) * np.log(net.config_ms.q.L) / conversion
    return sum(costs_bpsp), [float(c.detach()) for c in costs_bpsp] + [float(final)]


def make_optimizer(blueprint):
    """multiscale_trainer.py:72-80: optimiser class and initial learning rate from the model config."""
    cfg = blueprint.net.config_ms
    cls = {'RMSprop': torch.optim.RMSprop, 'Adam': torch.optim.Adam, 'SGD': torch.optim.SGD}[cfg.optim]
    return cls(blueprint.net.parameters(), cfg.lr.initial, weight_decay=cfg.weight_decay)


def train_step(blueprint, optimizer, img_batch):
    """One optimisation step on an NCHW float batch in [0, 255] (multiscale_trainer.py:173-200).
    -> (loss_pc as float, bpsp per scale)."""
    blueprint.net.zero_grad()
    loss_pc, bpsps = training_loss(blueprint, img_batch)
    loss_pc.backward()
    optimizer.step()
    return float(loss_pc.detach()), bpsps
