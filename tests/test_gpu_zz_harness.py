"""GPU: the test harness end to end on a throw-away experiment directory:
seed-0 weights saved as a reference-format checkpoint, three PNGs (two of equal size, one that needs
padding), default (theoretical bpsp, batched) mode against the one-image-at-a-time path, then
--write_to_files through the real coder.  (The file name sorts after the hot-path suites on purpose:
the harness is a caller of the path, its test runs once the path itself is green.)"""
import os

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _png(path, h, w, seed):
    from PIL import Image
    Image.fromarray(util.make_image(seed, h, w).permute(1, 2, 0).contiguous().numpy()).save(path)


def test_harness_theory_and_write_to_files(tmp_path, capsys):
    from l3c_pytorch_b200 import harness as H
    bp0 = util.blueprint('cr')
    exp = tmp_path / 'logs' / '0306_0001 cr oi'
    (exp / 'ckpts').mkdir(parents=True)
    torch.save({'net': {k: v.detach().cpu() for k, v in bp0.net.state_dict().items()}},
               str(exp / 'ckpts' / 'ckpt_0000000010.pt'))
    imgs = tmp_path / 'imgs'
    imgs.mkdir()
    _png(str(imgs / 'a.png'), 64, 64, 1)
    _png(str(imgs / 'b.png'), 64, 64, 2)
    _png(str(imgs / 'c.png'), 44, 36, 3)                              # padded to 48 x 40 inside
    logs = str(tmp_path / 'logs')

    flags = H.build_parser().parse_args([logs, '0306_0001', str(imgs)])
    tester = H.Tester('0306_0001', flags, -1)
    assert tester.restore_itr == 10
    res = tester.test(H.Testset(str(imgs)))
    assert sorted(res.per_img) == ['a', 'b', 'c']
    # the batched evaluation equals the reference-style evaluation of one image at a time
    fac = 2 ** tester.config_ms.num_scales
    for name in ('a', 'b', 'c'):
        raw = H.read_image_chw(str(imgs / (name + '.png'))).unsqueeze(0)
        batch, _ = tester.blueprint.unpack_batch_pad(raw, fac=fac)
        out = tester.blueprint.forward(batch)
        one = tester.blueprint.get_loss(out, num_subpixels_before_pad=int(np.prod(raw.shape)))
        want = float(sum(one.nonrecursive_bpsps))
        assert abs(res.per_img[name] - want) <= 1e-5 * want, (name, res.per_img[name], want)
    # THEORETICAL cost of uint8 noise under default-init weights: ~39 bpsp (the reference itself reports
    # 36.4 + 2.3 + 0.6 + 0.1 at 32^2, tests/golden/summary.json `ref_theory_bpsps`: the cross-entropy is
    # not floored at 16 bits per symbol the way the 16-bit coder is); c also pays for its zero padding
    # (48*40 coded sub-pixels counted against 44*36) -- measured 40.0: zeros are far cheaper than noise
    for name in ('a', 'b'):
        assert 36.0 < res.per_img[name] < 43.0, (name, res.per_img[name])
    assert 36.0 < res.per_img['c'] < 43.0 * (48 * 40) / (44 * 36), res.per_img['c']

    assert H.main([logs, '0306_0001', str(imgs), '--names', 'seed0']) == 0        # served from the cache
    out = capsys.readouterr().out
    assert '*** Found cached' in out and 'seed0 (0306_0001)' in out and 'bpsp=' in out

    out_dir, rep = str(tmp_path / 'l3c_out'), str(tmp_path / 'times.txt')
    assert H.main([logs, '0306_0001', str(imgs), '--write_to_files', out_dir, '--time_report', rep]) == 0
    assert sorted(os.listdir(out_dir)) == ['a.l3c', 'b.l3c', 'c.l3c']           # lossless: checked inside
    # CODED size: every RGB symbol costs at most 16 bits + the three bottleneck scales at most
    # log2(25) bits per symbol (5/4, 5/16, 5/64 symbols per pixel) + headers -> ~18.6 bpsp on the PADDED
    # sub-pixels (what the reference measures, bitcoding.py:108-110); far BELOW the theoretical cost above
    for name in ('a', 'b', 'c'):
        padded = 3 * (64 * 64 if name != 'c' else 48 * 40)
        real = 8.0 * os.path.getsize(os.path.join(out_dir, name + '.l3c')) / padded
        assert 15.0 < real < 19.2, (name, real)
        assert real < res.per_img[name]
    assert open(rep).read().startswith('Average times:')
