"""CPU: host logic of the test harness -- experiment / config / checkpoint
resolution, test sets, flags, and the tester's control flow with stand-in model classes (the real
classes need a GPU; their numerics are covered by tests/test_gpu_*.py)."""
import os

import numpy as np
import pytest
import torch

from l3c_pytorch_b200 import harness as H


def _png(path, h, w, seed):
    from PIL import Image
    rng = np.random.default_rng(seed)
    Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).save(path)


@pytest.fixture
def tree(tmp_path):
    """logs/<two experiments>, configs/{ms,dl}, imgs/ (3 PNGs, two of equal size)."""
    logs, cfgs, imgs = tmp_path / 'logs', tmp_path / 'configs', tmp_path / 'imgs'
    for d in (cfgs / 'ms', cfgs / 'dl', cfgs / 'ms' / 'sub', imgs):
        d.mkdir(parents=True)
    here = os.path.join(os.path.dirname(H.__file__), 'configs', 'ms', 'cr.cf')
    for name in ('cr.cf', 'lr1e-5.cf', 'lr1e-4.cf'):
        (cfgs / 'ms' / name).write_text(open(here).read())
    (cfgs / 'ms' / 'sub' / 'deep.cf').write_text(open(here).read())
    (cfgs / 'dl' / 'oi.cf').write_text('batchsize_train = 30\n')
    e1 = logs / '0306_0001 cr oi'
    e2 = logs / '0307_1200 sub@deep oi r@0306_0001 note'
    for e, itrs in ((e1, (1000, 2000, 3000)), (e2, (500,))):
        (e / 'ckpts').mkdir(parents=True)
        for it in itrs:
            torch.save({'net': {'w': torch.tensor([float(it)])}}, str(e / 'ckpts' / ('ckpt_%010d.pt' % it)))
    torch.save({'net': {'w': torch.tensor([3500.])}}, str(e1 / 'ckpts' / 'ckpt_0000003500.pt.tmp'))
    _png(str(imgs / 'a.png'), 16, 24, 1)
    _png(str(imgs / 'b.png'), 16, 24, 2)
    _png(str(imgs / 'c.png'), 8, 8, 3)
    (imgs / 'notes.txt').write_text('not an image')
    return dict(logs=str(logs), cfgs=str(cfgs), imgs=str(imgs), e1=str(e1), e2=str(e2), root=tmp_path)


def test_experiment_config_and_checkpoint_resolution(tree):
    assert H.is_log_date('0306_0001') and not H.is_log_date('306_0001') and not H.is_log_date('cr')
    assert H.find_experiment_dir(tree['logs'], '0306_0001') == tree['e1']
    with pytest.raises(ValueError):
        H.find_experiment_dir(tree['logs'], '0101_0000')
    (ms, dl), post = H.configs_of_experiment(tree['e1'], tree['cfgs'])
    assert ms.endswith(os.path.join('ms', 'cr.cf')) and dl.endswith(os.path.join('dl', 'oi.cf')) and post == ()
    (ms2, _), post2 = H.configs_of_experiment(tree['e2'], tree['cfgs'])          # '@' = path separator, r@ skipped
    assert ms2.endswith(os.path.join('ms', 'sub', 'deep.cf')) and post2 == ('note',)
    # '*' stands for exactly one character: lr1e*5 must find lr1e-5 only
    assert H._resolve_config(os.path.join(tree['cfgs'], 'ms'), 'lr1e*5').endswith('lr1e-5.cf')
    with pytest.raises(ValueError):
        H._resolve_config(os.path.join(tree['cfgs'], 'ms'), 'lr1e*')              # one char too short
    with pytest.raises(ValueError):
        H._resolve_config(os.path.join(tree['cfgs'], 'ms'), 'nope')
    (ms3, dl3), _ = H.configs_of_experiment(tree['e1'], H.default_configs_dir())   # package tree: ms/ only
    assert ms3.endswith('cr.cf') and dl3 is None
    ck = H.list_checkpoints(tree['e1'])
    assert [i for i, _ in ck] == [1000, 2000, 3000, 3500]
    assert H.checkpoint_for_itr(ck, -1)[0] == 3500                                # newest, temporary ones count
    assert H.checkpoint_for_itr(ck, 2000)[0] == 2000 and H.checkpoint_for_itr(ck, 2999)[0] == 2000
    with pytest.raises(ValueError):
        H.checkpoint_for_itr(ck, 10)


def test_testsets_and_flags(tree):
    ts = H.Testset(tree['imgs'])
    assert len(ts) == 3 and ts.id == 'imgs_3' and all(p.endswith('.png') for p in ts.ps)
    assert H.Testset(tree['imgs'], max_imgs=2).id == 'imgs_2'
    ts.filter_filenames(['a', 'c'])
    assert [os.path.basename(p) for p in ts.ps] == ['a.png', 'c.png']
    with pytest.raises(ValueError):
        ts.filter_filenames(['zzz'])
    one = H.Testset(os.path.join(tree['imgs'], 'b.png'), append_id='_crop8')
    assert len(one) == 1 and one.id.endswith('b.png_crop8')
    with pytest.raises(FileNotFoundError):
        H.Testset(os.path.join(tree['imgs'], 'missing.png'))
    img = H.read_image_chw(os.path.join(tree['imgs'], 'a.png'), crop=8)
    full = H.read_image_chw(os.path.join(tree['imgs'], 'a.png'))
    assert img.shape == (3, 8, 8) and img.dtype == torch.uint8 and torch.equal(img, full[:, 4:12, 8:16])
    p = H.build_parser()
    for argv, msg in [(['l', '0306_0001', 'i', '--compare_theory'], 'compare_theory'),
                      (['l', '0306_0001', 'i', '--time_report', 't'], 'time_report'),
                      (['l', '0306_0001', 'i', '--write_to_files', 'o', '--sample', 's'], 'sample')]:
        with pytest.raises(ValueError, match=msg):
            H.check_flags(p.parse_args(argv))
    assert H.format_table([('Testset', 'Itr'), ('imgs_3', '7')]) == 'Testset  Itr\nimgs_3   7'


class _FakeNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))


class _FakeBlueprint(object):
    """bpsp of an image := mean pixel value / 255 + w / 1e6 (w comes from the checkpoint)."""
    made = []

    def __init__(self, config_ms):
        self.config_ms, self.net, self.device = config_ms, _FakeNet(), torch.device('cpu')
        self.batches = []
        _FakeBlueprint.made.append(self)

    def set_eval(self):
        return self

    def unpack_batch_pad(self, raw, fac):
        assert raw.dim() == 4 and raw.dtype == torch.uint8
        return raw.float(), raw.long()

    def forward(self, img_batch, auto_recurse=0):
        self.batches.append(tuple(img_batch.shape))
        self.recursions = getattr(self, 'recursions', []) + [auto_recurse]
        return img_batch

    def get_loss(self, out):
        class L(object):
            nonrecursive_bpsps = [3.0, 1.0, 0.5, 0.25]
        return L()

    def sample_forward(self, img_batch, sample_scales, partial_final=None):
        assert sample_scales in ([], [0], [0, 1])
        return img_batch.float() * 0.5

    def get_loss_per_image(self, out, num_subpixels_before_pad=None):
        assert num_subpixels_before_pad == int(np.prod(out.shape[1:]))
        return [float(x.mean() / 255. + self.net.w.item() / 1e6) for x in out]


class _FakeBitcoding(object):
    def __init__(self, blueprint, times, compare_with_theory=False):
        self.times, self.compare_with_theory = times, compare_with_theory

    def encode(self, img, pout):
        assert not os.path.isfile(pout) and img.dtype == torch.int64 and img.shape[0] == 1
        np.save(pout + '.npy', img.numpy())
        os.rename(pout + '.npy', pout)
        return 8.0 * os.path.getsize(pout) / img.numel()

    def decode(self, pin):
        return torch.from_numpy(np.load(open(pin, 'rb')))


def _fake_tester(log_date, flags, restore_itr, configs_dir=None):
    return H.Tester(log_date, flags, restore_itr, configs_dir=configs_dir,
                    make_blueprint=_FakeBlueprint, make_bitcoding=_FakeBitcoding)


def test_tester_control_flow_with_stand_ins(tree, capsys):
    _FakeBlueprint.made.clear()
    argv = [tree['logs'], '0306_0001', tree['imgs'], '--configs_dir', tree['cfgs'], '--restore_itr', '2500',
            '--names', 'mine', '--batch', '4']
    assert H.main(argv, tester_cls=_fake_tester) == 0
    bp = _FakeBlueprint.made[-1]
    assert bp.net.w.item() == 2000.0                                    # closest checkpoint not after 2500
    assert sorted(bp.batches) == [(1, 3, 8, 8), (2, 3, 16, 24)]         # equal-sized images share a pass
    out = capsys.readouterr().out
    assert '*** Summary:' in out and 'mine (0306_0001)' in out and 'imgs_3' in out and 'bpsp=' in out
    want = np.mean([float(H.read_image_chw(os.path.join(tree['imgs'], n)).float().mean() / 255. + 2000 / 1e6)
                    for n in ('a.png', 'b.png', 'c.png')])
    got = float(out.split('bpsp=')[-1].split()[0])
    assert abs(got - want) < 1e-9
    # second run: served from LOG_DIR_test/<experiment>/results.json, no forward pass
    assert H.main(argv, tester_cls=_fake_tester) == 0
    assert _FakeBlueprint.made[-1].batches == [] and '*** Found cached' in capsys.readouterr().out
    assert os.path.isfile(os.path.join(tree['logs'] + '_test', '0306_0001 cr oi', 'results.json'))
    # --write_to_files: real files, decoded and compared, time report written, no summary
    out_dir, rep = str(tree['root'] / 'l3c_out'), str(tree['root'] / 'times.txt')
    assert H.main([tree['logs'], '0306_0001', tree['imgs'], '--configs_dir', tree['cfgs'], '--write_to_files', out_dir,
                   '--compare_theory', '--time_report', rep], tester_cls=_fake_tester) == 0
    assert sorted(os.listdir(out_dir)) == ['a.l3c', 'b.l3c', 'c.l3c']
    txt = open(rep).read()
    assert txt.startswith('Average times:') and '=== bc.encode' in txt and '=== bc.decode' in txt
    assert '*** Summary:' not in capsys.readouterr().out
    # --recursive: ignored for L3C proper (multiscale_tester.py:123-125), not combinable with --write_to_files
    assert H.main(argv + ['--recursive', 'auto', '--overwrite_cache'], tester_cls=_fake_tester) == 0
    assert set(_FakeBlueprint.made[-1].recursions) == {0}
    # --sample: ground truth + three sampled images per test image, refuses to overwrite
    sdir = str(tree['root'] / 'samples')
    assert H.main(argv + ['--sample', sdir, '--overwrite_cache'], tester_cls=_fake_tester) == 0
    names = sorted(os.listdir(os.path.join(sdir, '0306_0001')))
    assert len(names) == 12 and sum(n.endswith('_gt.png') for n in names) == 3
    assert any('_rgb+bn0+bn1_' in n for n in names) and any('_rgb_' in n for n in names)
    with pytest.raises(FileExistsError):
        H.main(argv + ['--sample', sdir, '--overwrite_cache'], tester_cls=_fake_tester)


def test_recursive_flag():
    """multiscale_tester.py:123-132"""
    class C(object):
        def __init__(self, rgb, n):
            self.rgb_bicubic_baseline, self.num_scales = rgb, n
    assert H.parse_recursive_flag('auto', C(False, 3)) == 0 and H.parse_recursive_flag('2', C(False, 3)) == 0
    assert H.parse_recursive_flag('auto', C(True, 1)) == 3 and H.parse_recursive_flag('2', C(True, 1)) == 2
    assert H.parse_recursive_flag('auto', C(True, 3)) == 0 and H.parse_recursive_flag('0', C(True, 1)) == 0
