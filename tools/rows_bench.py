"""Micro-benchmark of the CDF-row builder (dmll_table_kernel) on the real parameter tensor of the bench batch:
16 x 3x512x512 noise images, seed-0 default-init L3C weights, f16 conv mode.  Prints ms per full channel plane
(786 432 sub-pixels x 16 images x 256 entries x 10 mixture terms) for R, G, B and for the bottleneck scales."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import l3c_pytorch_b200 as l3c
    from l3c_pytorch_b200 import config, engine as E
    E.set_conv_precision('f16')
    torch.manual_seed(0)
    bp = l3c.MultiscaleBlueprint(config.ms_config('cr')).set_eval()
    imgs = torch.stack([(torch.rand(3, 512, 512, generator=torch.Generator().manual_seed(1000 + i)) * 255).round()
                        .to(torch.uint8) for i in range(16)]).cuda()
    out = bp.net(imgs)
    K = 10
    for scale in (0, 1, 2):
        l, S = out.P_nhwc[scale], out.S_u8[scale]
        N, C, H, W = S.shape
        dm = bp.losses.loss_dmol_rgb if scale == 0 else bp.losses.loss_dmol_n
        tg = dm.targets(l.device)
        pitch = E.table_pitch(dm.L)
        table = torch.empty(N * C * H * W * pitch, dtype=torch.int16, device=l.device)
        for c in (range(C) if dm.rgb_scale else [-1]):
            for _ in range(2):
                E.dmll_build_table(l, S, tg, C, K, dm.L, dm.rgb_scale, c, table)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                E.dmll_build_table(l, S, tg, C, K, dm.L, dm.rgb_scale, c, table)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            rows = N * H * W * (1 if c >= 0 else C)
            print('scale %d channel %2d: %.3f ms  (%d rows, %.1f GB/s of rows written)'
                  % (scale, c, ms, rows, rows * pitch * 2 / ms / 1e6), flush=True)


if __name__ == '__main__':
    main()
