"""CPU: host-side logic of the product (no kernels launched) + the C-ABI surface."""
import ctypes
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from tests import util

ROOT = util.ROOT


def _digest(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


@pytest.mark.parametrize('cfg,key', [('cr', 'l3c'), ('cr_rgb_shared', 'rgbs')])
def test_state_dict_matches_reference(cfg, key):
    """same keys, shapes, order and seed-0 default-init VALUES as the reference's module tree."""
    bp = util.blueprint(cfg, device='cpu')
    sd = bp.net.state_dict()
    g = util.golden_summary()
    assert sum(v.numel() for v in sd.values()) == g[key + '_sd_numel']
    assert _digest(sd) == g[key + '_sd_sha256']
    if key == 'l3c':
        assert len(sd) == 268
        for k, probe in g['l3c_sd_probe'].items():
            np.testing.assert_allclose(sd[k].flatten()[:3].numpy(), probe, rtol=0, atol=0)


def test_library_exports_every_declared_symbol():
    from l3c_pytorch_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'l3c_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(l3c_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(lib, name), 'header declares %s but the library does not export it' % name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    # the library is sm_100a code
    out = subprocess.run(['cuobjdump', '-lelf', _lib.SO_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_uniform_row_host_function_matches_oracle():
    from l3c_pytorch_b200 import engine
    from oracle import ac
    for L in (25, 256, 5):
        assert (engine.uniform_cdf_row(L) == ac.uniform_cdf_row(L)).all()


def test_config_parser():
    from l3c_pytorch_b200 import config
    c = config.ms_config('cr')
    assert (c.num_scales, c.Cf, c.q.C, c.q.L, c.prob.K, c.enc.num_blocks) == (3, 64, 5, 25, 10, 8)
    assert c.q.levels_range == (-1, 1) and c.dec.skip is True and c.rgb_bicubic_baseline is False
    s = config.ms_config('cr_rgb_shared')
    assert (s.num_scales, s.q.C, s.enc.cls, s.dec.skip, s.enc.feed_F) == (1, 3, 'BicubicSubsampling', False, False)
    assert s.Cf == 64 and s.rgb_bicubic_baseline is True            # inherited through `use cr.cf`
    assert config.ms_config('cr_rgb').num_scales == 3


def test_pad_matches_reference_rule():
    from l3c_pytorch_b200 import pad
    from oracle import model as om
    for h, w in [(32, 32), (40, 28), (50, 30), (1500, 1000), (7, 9), (8, 15)]:
        assert pad.padding_tuple(h, w, 8) == om.pad_tuple(h, w, 8)
    x = torch.arange(3 * 50 * 30).reshape(1, 3, 50, 30)
    y, t = pad.pad(x, 8, 'constant')
    assert y.shape == (1, 3, 56, 32) and t == (1, 1, 3, 3)
    assert (pad.undo_pad(y, *t) == x).all()
    assert pad.padding_tuple(1500, 1000, 8) == (0, 0, 2, 2)          # 1504x1000


def test_auto_crop_counts_and_stitch(monkeypatch):
    from l3c_pytorch_b200 import auto_crop
    for H, W, n in [(1000, 600, 64), (492, 326, 16), (204, 204, 4), (102, 102, 1)]:
        img = (torch.rand(1, 3, H, W) * 255).round().long()
        crops = list(auto_crop.iter_crops(img, 204 * 102))
        assert len(crops) == n
        if n > 1:
            assert (auto_crop.stitch(crops) == img).all()
    monkeypatch.setenv('AC_NEEDS_CROP_DIM', '40,40')
    img = torch.zeros(1, 3, 100, 60)
    assert auto_crop.needs_crop(img) and len(list(auto_crop.iter_crops(img))) == 4
    monkeypatch.delenv('AC_NEEDS_CROP_DIM')
    assert not auto_crop.needs_crop(torch.zeros(1, 3, 1500, 2000))
    assert auto_crop.needs_crop(torch.zeros(1, 3, 3000, 2000))
    c = auto_crop.CropLossCombinator()
    c.add(2.0, 100)
    c.add(4.0, 300)
    assert abs(c.get_bpsp() - 3.5) < 1e-12


def test_part_suffix(tmp_path):
    from l3c_pytorch_b200 import part_suffix_helper as ps
    assert ps.make_part_suffix(10) == '.part10'
    assert ps.contains_part_suffix('a/b.part3') and not ps.contains_part_suffix('a/b.part3/more')
    assert ps.index_of_part_suffix('a/b.part13') == 13
    for i in range(16):
        (tmp_path / ('some.file.part%d' % i)).write_text('x')
    (tmp_path / 'some.file.partX').write_text('x')
    got = ps.iter_part_suffixes(str(tmp_path / 'some.file.part3'))
    assert [os.path.basename(p) for p in got] == ['some.file.part%d' % i for i in range(16)]


def test_container_layout_roundtrip_and_reference_header():
    from l3c_pytorch_b200.codec import ContainerLayout, parse_container, MAGIC
    shapes = [(3, 5, 4, 4), (2, 5, 8, 8), (1, 5, 16, 16), (0, 3, 32, 32)]
    lens = list(range(3, 21))
    lay = ContainerLayout(shapes)
    total, pieces, offs = lay.header_and_offsets(lens, (1, 2, 3, 4))
    buf = bytearray(total)
    for o, b in pieces:
        buf[o:o + len(b)] = b
    for o, n in zip(offs, lens):
        buf[o:o + n] = bytes([n]) * n
    pt, scales = parse_container(bytes(buf))
    assert pt == (1, 2, 3, 4)
    assert [(C, H, W) for (C, H, W, _) in scales] == [(5, 4, 4), (5, 8, 8), (5, 16, 16), (3, 32, 32)]
    assert [n for s in scales for (_, n) in s[3]] == lens
    assert total == 8 + 4 * (5 + 4) + 4 * 18 + sum(lens)
    # a real file written by the reference parses, and its header equals what we would write
    g = util.golden_npz('l3c_40x28_i1')
    data = g['container'].tobytes()
    pt, scales = parse_container(data)
    assert pt == tuple(util.golden_summary()['l3c_40x28_i1']['pad'])
    lens = [n for s in scales for (_, n) in s[3]]
    shapes = [(3 - i, C, H, W) for i, (C, H, W, _) in enumerate(scales)]
    total, pieces, offs = ContainerLayout(shapes).header_and_offsets(lens, pt)
    assert total == len(data)
    for o, b in pieces:
        assert data[o:o + len(b)] == b
    with pytest.raises(ValueError):
        parse_container(data[:-1])
    with pytest.raises(ValueError):
        parse_container(data[:100])
    assert MAGIC == bytes([0x46, 0xE2, 0x84, 0x92])


def test_header_field_helpers(tmp_path):
    from l3c_pytorch_b200 import bitcoding as bc
    p = tmp_path / 'x.l3c'
    with open(p, 'wb') as f:
        bc.write_padding_tuple((1, 2, 3, 4), f)
        bc.write_shape((1, 3, 512, 768), f)
        bc.write_num_bytes_encoded(1234567, f)
    with open(p, 'rb') as f:
        assert bc.read_padding_tuple(f) == (1, 2, 3, 4)
        assert bc.read_shapes(f) == (3, 512, 768)
        assert bc.read_num_bytes_encoded(f) == 1234567
    assert open(p, 'rb').read().hex() == '0100020003000400' + '03' + '0002' + '0003' + '87d61200'


def test_torchac_shim_argument_errors():
    """same exception types as the reference shim (torchac.py:93-94,127-131)."""
    from l3c_pytorch_b200 import torchac
    cdf = torch.zeros(1, 1, 2, 26, dtype=torch.int16)
    if not torch.cuda.is_available():
        with pytest.raises(ValueError):
            torchac.encode_cdf(cdf, torch.zeros(2, dtype=torch.int16))      # no CPU backend: loud
    t = torch.zeros(26)
    with pytest.raises(ValueError):
        torchac.encode_logistic_mixture(t, torch.zeros(1, 2, 1, 2), torch.zeros(1, 2, 1, 2),
                                        torch.zeros(1, 2, 1, 2), torch.zeros(2, dtype=torch.int16))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'l3c_pytorch_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), fn
                assert 'liboracle' not in src and 'oracle/' not in src.replace('oracle/ is', ''), fn


def test_no_cpu_fallback_on_cpu_tensors():
    from l3c_pytorch_b200 import engine
    bp = util.blueprint('cr', device='cpu')
    with pytest.raises(ValueError):
        bp.net(torch.zeros(1, 3, 32, 32))
    with pytest.raises(ValueError):
        engine.conv2d(bp.net.heads[1].head, torch.zeros(1, 8, 8, 64))


def test_coder_state_and_batched_emission_models_match_the_oracle():
    """The (low, r) state with the collapsed renormalisation and the order-free, per-32-symbol bit
    emission of ac_encode_kernel, modelled in Python, reproduce the oracle's (= the reference's) bytes --
    including owed-bit runs that cross batches and exceed 32 (the kernel's sequential-sink case)."""
    from oracle import ac
    from tests import kernel_models as km
    rng = np.random.default_rng(11)
    cases = []
    for L, n in [(25, 1), (25, 200), (256, 333), (3, 97)]:
        w = rng.integers(1, 4000, size=(n, L)).astype(np.float64)
        w = w / w.sum(1, keepdims=True) * (65536 - L - 41)
        c = np.floor(np.cumsum(w, 1)).astype(np.int64) + np.arange(1, L + 1)
        cdf = np.concatenate([np.zeros((n, 1), np.int64), c[:, :-1], np.zeros((n, 1), np.int64)], 1).astype(np.uint16)
        cases.append((cdf, rng.integers(0, L, size=n).astype(np.int16)))
    for n in (5, 33, 64, 400):                                   # long underflow runs
        rows = np.zeros((n, 4), np.int64)
        for i in range(n):
            rows[i, :3] = [0, 32768 - int(rng.integers(1, 200)), 32768 + int(rng.integers(1, 200))]
        sym = np.ones(n, np.int16)
        sym[rng.integers(0, n, size=max(1, n // 37))] = 0
        cases.append((rows.astype(np.uint16), sym))
        cases.append((rows.astype(np.uint16), np.ones(n, np.int16)))
    for cdf, sym in cases:
        L = cdf.shape[1] - 1
        iv = []
        for i, s_ in enumerate(sym):
            hi = 65536 if s_ == L - 1 else int(cdf[i, s_ + 1])
            iv.append((int(cdf[i, s_]), hi))
        recs, final_low = km.coder_records(iv)
        data = ac.encode(cdf, sym)
        assert km.emit_batched(recs, final_low) == data
        # and the decoders' (low, r, dv) state with the division-free search reads it back
        assert (km.decode_model(cdf, data, len(sym)) == sym).all()
        # garbage / truncated input: same symbols as the oracle (= the compiled reference, pin_oracle.py)
        junk = bytes(rng.integers(0, 256, size=max(1, len(data) // 2)).astype(np.uint8))
        assert (km.decode_model(cdf, junk, len(sym)) == ac.decode(cdf, junk)).all()


def test_container_shapes_are_validated_before_any_buffer_is_sized():
    """C/H/W of every scale come from the file; a crafted or wrong-config container must raise
    ValueError (the reference fails with a torch shape error) instead of reaching the kernels."""
    from l3c_pytorch_b200.codec import BatchCodec
    codec = BatchCodec(util.blueprint('cr', device='cpu'))
    good = [(5, 4, 4), (5, 8, 8), (5, 16, 16), (3, 32, 32)]
    codec._check_shapes(good)
    for bad in ([(5, 4, 4), (8, 8, 8), (5, 16, 16), (3, 32, 32)],      # C=8 at a q.C=5 scale
                [(5, 4, 4), (5, 8, 8), (5, 16, 16), (3, 16, 16)],      # H/W smaller than the net's output
                [(3, 4, 4), (5, 8, 8), (5, 16, 16), (3, 32, 32)],      # wrong C at the uniform scale
                [(5, 4, 4), (5, 8, 9), (5, 16, 18), (3, 32, 36)],      # not a factor of 2
                [(5, 0, 4), (5, 0, 8), (5, 0, 16), (3, 0, 32)]):
        with pytest.raises(ValueError):
            codec._check_shapes(bad)
    rgb = BatchCodec(util.blueprint('cr_rgb_shared', device='cpu'))
    rgb._check_shapes([(3, 32, 32), (3, 64, 64)])
    with pytest.raises(ValueError):
        rgb._check_shapes([(5, 32, 32), (3, 64, 64)])


def test_part_files_are_never_overwritten(tmp_path):
    """the reference asserts on every part path (bitcoding.py:57 through the recursion of
    :63-71); stale higher-numbered parts must not survive a new encode either."""
    from l3c_pytorch_b200 import part_suffix_helper as psh
    p = str(tmp_path / 'x.l3c')
    assert psh.existing_parts(p) == []
    for i in (0, 1, 10, 2):
        open(p + psh.make_part_suffix(i), 'wb').close()
    open(p + '.partial', 'wb').close()                               # not a part file
    assert [psh.index_of_part_suffix(q) for q in psh.existing_parts(p)] == [0, 1, 2, 10]


def test_tiled_container_layout_and_parser_round_trip():
    """Tiled layout (codec.ContainerLayout with tile=(th, tw)): tile geometry incl. ragged edges, header tag,
    parse_container recovers every stream; the reference layout is untouched."""
    import struct
    import numpy as np
    from l3c_pytorch_b200 import codec as C

    def order_to_raster(r, H, W, th, tw):          # Python restatement of dmll.cu: tile_order_to_raster
        ty = r // (th * W)
        h = min(th, H - ty * th)
        rem = r - ty * th * W
        tx = rem // (tw * h)
        w = min(tw, W - tx * tw)
        rem2 = rem - tx * tw * h
        return (ty * th + rem2 // w) * W + tx * tw + rem2 % w

    for (H, W, tile) in [(64, 96, (16, 16)), (37, 50, (16, 24)), (8, 8, (64, 64)), (5, 130, (2, 64))]:
        grid = C.tile_grid(H, W, tile)
        assert sum(n for (_, n) in grid) == H * W and grid[0][0] == 0
        assert all(grid[i][0] + grid[i][1] == grid[i + 1][0] for i in range(len(grid) - 1))
        seen = np.array([order_to_raster(r, H, W, *tile) for r in range(H * W)])
        assert sorted(seen.tolist()) == list(range(H * W))                     # a permutation
        # the first stream is the top-left tile, row-major inside
        h0, w0 = min(tile[0], H), min(tile[1], W)
        assert seen[:h0 * w0].tolist() == [y * W + x for y in range(h0) for x in range(w0)]
    assert C.tile_grid(7, 9, None) == [(0, 63)]

    shapes = [(3, 5, 8, 12), (2, 5, 16, 24), (1, 5, 32, 48), (0, 3, 64, 96)]
    for tile in (None, (16, 16), (32, 40)):
        layout = C.ContainerLayout(shapes, tile)
        n_streams = sum(Cc * len(C.tile_grid(H, W, tile)) for (_, Cc, H, W) in shapes)
        rng = np.random.default_rng(1)
        lens = rng.integers(0, 40, n_streams)
        total, pieces, offs = layout.header_and_offsets(lens, (1, 2, 3, 4))
        buf = bytearray(total)
        for (o, b) in pieces:
            buf[o:o + len(b)] = b
        for i, (o, n) in enumerate(zip(offs, lens)):
            buf[o:o + n] = bytes([i % 251]) * int(n)
        data = bytes(buf)
        assert C.container_tile(data) == tile
        pt, scales = C.parse_container(data)
        assert pt == (1, 2, 3, 4) and [(c, h, w) for (c, h, w, _) in scales] == [(c, h, w) for (_, c, h, w) in shapes]
        flat = [st for (_, _, _, sts) in scales for st in sts]
        assert [o for (o, _) in flat] == list(offs) and [n for (_, n) in flat] == [int(x) for x in lens]
        if tile is None:
            assert data[8] == 5 and struct.unpack_from('<HH', data, 9) == (8, 12)      # reference layout untouched


def test_f16x2_split_weight_images():
    """host side of precision mode 'f16x2' (engine.PackedConv.get_f16x2 uses engine._split_host): hi + lo / 2^11
    reproduces an fp32 weight to 2^-21 relative, the images have the layout conv_f16x2.cu's TMA maps read
    (3x3: [9][cout_pad][hi 64 | lo 64], 1x1: [hi chunks..., lo chunks...][cout_pad][64]), padding rows are zero."""
    import numpy as np
    import torch
    from l3c_pytorch_b200 import engine as E
    torch.manual_seed(3)
    w = torch.randn(1000) * torch.logspace(-6, 3, 1000)
    hi, lo = E._split_host(w)
    assert hi.dtype == torch.float16 and lo.dtype == torch.float16
    back = hi.double() + lo.double() / 2048.0
    big = w.abs() > 1e-3                     # above the FP16 subnormal range the split carries 22 bits
    assert float(((back - w.double()).abs() / w.double().abs())[big].max()) < 2.0 ** -21
    assert float((back - w.double()).abs()[~big].max()) < 1e-9

    class FakePacked(E.PackedConv):
        def get(self):                       # no device repack on a CPU box: only the split image is under test
            self._key = 'k'
            cout = self.conv.weight.shape[0]
            return None, torch.zeros((cout + 63) // 64 * 64)

    for k, cin, cout in ((3, 64, 64), (3, 64, 256), (1, 192, 120), (1, 64, 10)):
        conv = torch.nn.Conv2d(cin, cout, k)
        img, b = FakePacked(conv).get_f16x2()
        cp = b.shape[0]
        wd = conv.weight.detach()
        if k == 3:
            assert tuple(img.shape) == (9, cp, 128)
            val = img[:, :, :64].double() + img[:, :, 64:].double() / 2048.0          # [tap][cout][cin]
            want = wd.permute(2, 3, 0, 1).reshape(9, cout, 64).double()
        else:
            nch = cin // 64
            assert tuple(img.shape) == (2 * nch, cp, 64)
            val = img[:nch].double() + img[nch:].double() / 2048.0                    # [chunk][cout][64]
            want = wd.reshape(cout, nch, 64).permute(1, 0, 2).double()
        np.testing.assert_allclose(val[:, :cout].numpy(), want.numpy(), rtol=2.0 ** -20, atol=1e-9)
        assert bool((img[:, cout:] == 0).all())


def test_f16x2_split_product_model():
    """Numerics of precision mode 'f16x2' in numpy (what conv_f16x2.cu's three MMAs compute): with
    x = hi_x + lo_x / 2^11 and w = hi_w + lo_w / 2^11 (all four FP16), a K = 576 dot product evaluated as
    sum(hi_x hi_w) + sum(hi_x lo_w + lo_x hi_w) / 2^11 with fp32 accumulation is within a few fp32 roundings of the
    float64 result -- ~1000x closer than the one-MMA FP16-operand mode, whose error is the operand rounding."""
    import numpy as np
    rng = np.random.default_rng(0)
    K, n = 576, 4000
    x = rng.standard_normal((n, K)).astype(np.float32)
    w = (rng.uniform(-1, 1, (n, K)) / 24).astype(np.float32)
    want = (x.astype(np.float64) * w.astype(np.float64)).sum(1)

    def split(a):
        hi = a.astype(np.float16)
        lo = ((a - hi.astype(np.float32)) * np.float32(2048)).astype(np.float16)
        return hi.astype(np.float32), lo.astype(np.float32)

    def acc32(p):                      # fp32 accumulation in K order (the tensor core's order is not specified:
        s = np.zeros(p.shape[0], np.float32)          # any order stays inside the bound asserted below)
        for k in range(p.shape[1]):
            s = s + p[:, k]
        return s

    hx, lx = split(x)
    hw, lw = split(w)
    acc1 = acc32(hx * hw)                                                   # products of FP16 pairs are exact in fp32
    acc2 = acc32(hx * lw) + acc32(lx * hw)
    got = acc1 + acc2 * np.float32(1.0 / 2048)
    one = acc32(hx * hw)                                                    # the fast mode: hi parts only
    f32 = acc32(x * w)
    scale = np.abs(x.astype(np.float64) * w).sum(1)                         # condition-free error scale
    e_split = np.abs(got - want) / scale
    e_one = np.abs(one - want) / scale
    e_f32 = np.abs(f32 - want) / scale
    assert e_split.max() < 4e-7, e_split.max()
    assert e_split.max() < 6 * e_f32.max()                                  # fp32-class
    assert np.median(e_one) > 200 * np.median(e_split)
