"""oracle/weights.py -- TEST INFRASTRUCTURE ONLY (also used by `bench.py --impl reference`).

Seed-0 default-init weights of the reference's module tree, built from plain `torch.nn` modules and
nothing of the product package: the reference arm of bench.py must not import l3c_pytorch_b200 (so
that the only native code mapped into that process is the checker's).

The tree below registers the same parameters, under the same names, in the same construction order
as /root/reference/src/modules/{multiscale_network.py:168-224, net.py:89-171, edsr.py:52-119,
head.py:26-59, prob_clf.py:29-74}, so `torch.manual_seed(0); build(cfg).state_dict()` consumes the RNG
exactly like `MultiscaleBlueprint(config)` does there: same keys, same order, same VALUES
(tests/test_oracle_kat.py checks the sha256 recorded from the unmodified reference in
tests/golden/summary.json).
"""
import torch
from torch import nn

RGB_MEAN = (0.4488, 0.4371, 0.4040)


def _conv(cin, cout, k, rate=1, stride=1):
    return nn.Conv2d(cin, cout, k, stride=stride, dilation=rate, padding=k // 2 if rate == 1 else rate)


class _MeanShift(nn.Conv2d):                                   # edsr.py:52-60
    def __init__(self, rgb_range, rgb_mean, rgb_std, sign=-1):
        super().__init__(3, 3, kernel_size=1)
        std = torch.Tensor(rgb_std)
        self.weight.data = torch.eye(3).view(3, 3, 1, 1)
        self.weight.data.div_(std.view(3, 1, 1, 1))
        self.bias.data = sign * rgb_range * torch.Tensor(rgb_mean)
        self.bias.data.div_(std)


class _ResBlock(nn.Module):                                    # edsr.py:63-89
    def __init__(self, cf):
        super().__init__()
        self.body = nn.Sequential(_conv(cf, cf, 3), nn.ReLU(True), _conv(cf, cf, 3))


def _body(cf, n_blocks):
    return nn.Sequential(*([_ResBlock(cf) for _ in range(n_blocks)] + [_conv(cf, cf, 3)]))


class _Head(nn.Module):                                        # head.py:41-59
    def __init__(self, cin, cf):
        super().__init__()
        self.head = _conv(cin, cf, 3)


class _RGBHead(nn.Module):                                     # head.py:26-38
    def __init__(self, cf):
        super().__init__()
        self.head = nn.Sequential(_MeanShift(0, (0., 0., 0.), (128., 128., 128.)), _Head(3, cf))


class _Q(nn.Module):                                           # quantizer.py:50-60
    def __init__(self, levels):
        super().__init__()
        self.levels = levels


class _Enc(nn.Module):                                         # net.py:89-127
    def __init__(self, cfg):
        super().__init__()
        self.down = _conv(cfg.Cf, cfg.Cf, 5, stride=2)
        self.body = _body(cfg.Cf, cfg.n_blocks_enc)
        self.to_q = nn.Sequential(_conv(cfg.Cf, cfg.qC, 1), nn.Identity())
        lo, hi = cfg.levels_range
        self.levels = nn.Parameter(torch.linspace(lo, hi, cfg.qL), requires_grad=False)
        self.q = _Q(self.levels)


class _Dec(nn.Module):                                         # net.py:151-171
    def __init__(self, cfg):
        super().__init__()
        self.head = _conv(cfg.qC, cfg.Cf, 1)
        self.body = _body(cfg.Cf, cfg.n_blocks_dec)
        self.tail = nn.Sequential(_conv(cfg.Cf, 4 * cfg.Cf, 3), nn.PixelShuffle(2))


class _Net(nn.Module):                                         # net.py:49-62
    def __init__(self, cfg):
        super().__init__()
        self.enc = nn.Module() if cfg.rgb_baseline else _Enc(cfg)       # BicubicDownsamplingEnc has no parameters
        self.dec = _Dec(cfg)


class _Atrous(nn.Module):                                      # prob_clf.py:44-74
    def __init__(self, cf, cout):
        super().__init__()
        self.atrous = nn.ModuleList([_conv(cf, cf, 3, rate=r) for r in (1, 2, 4)])
        self.lin = _conv(3 * cf, cout, 1)


class _ProbClf(nn.Module):                                     # prob_clf.py:29-41
    def __init__(self, cf, K, C):
        super().__init__()
        self.atrous = _Atrous(cf, (4 if C == 3 else 3) * C * K)


class _Network(nn.Module):                                     # multiscale_network.py:168-224
    def __init__(self, cfg):
        super().__init__()
        self.sub_rgb_mean = _MeanShift(255., RGB_MEAN, (1.0, 1.0, 1.0))
        S = cfg.num_scales
        if not cfg.rgb_baseline:
            heads = [_RGBHead(cfg.Cf)] + [_Head(cfg.Cf, cfg.Cf) for _ in range(S - 1)]
            nets = [_Net(cfg) for _ in range(S)]
            clfs = [_ProbClf(cfg.Cf, cfg.K, 3)] + [_ProbClf(cfg.Cf, cfg.K, cfg.qC) for _ in range(S - 1)]
        else:
            heads = [nn.Identity() for _ in range(S)]
            nets = [_Net(cfg) for _ in range(S)]
            clfs = [_ProbClf(cfg.Cf, cfg.K, 3) for _ in range(S)]
        self.heads, self.nets, self.prob_clfs = nn.ModuleList(heads), nn.ModuleList(nets), nn.ModuleList(clfs)


def default_init_state_dict(cfg, seed=0):
    """cfg: oracle.model.CFG_L3C / CFG_RGB_SHARED -> CPU state dict of the seeded default init."""
    torch.manual_seed(seed)
    net = _Network(cfg)
    return {k: v.detach().clone() for k, v in net.state_dict().items()}
