"""GPU: the whole path -- network, DMLL head, coder, container, crops -- against the oracle, the
committed goldens of the reference, and size-independent properties (lossless round trip)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

# fp32 FFMA path vs MKL-DNN fp32: only accumulation order differs, over ~40 layers.
# tf32 (tcgen05) path: operands rounded to 11 significant bits per conv -> looser.
TOL = {'fp32': dict(p=2e-3, flips=2e-3, bytes_rel=0.0), 'tf32': dict(p=8e-2, flips=2e-2, bytes_rel=2e-3),
       'f16': dict(p=8e-2, flips=2e-2, bytes_rel=2e-3),
       # strict tensor-core mode (split FP16 operands, 22 significant bits): fp32-class
       'f16x2': dict(p=4e-3, flips=4e-3, bytes_rel=5e-4)}


@pytest.fixture(params=['fp32', 'tf32', 'f16', 'f16x2'])
def prec(request):
    from l3c_pytorch_b200 import engine as E
    old = E.get_conv_precision()
    E.set_conv_precision(request.param)
    yield request.param
    E.set_conv_precision(old)


def _flip_stats(a, b):
    return float((a != b).float().mean())


@pytest.mark.parametrize('cfg,H,W', [('cr', 32, 32), ('cr', 64, 96), ('cr_rgb_shared', 64, 64)])
def test_forward_matches_oracle(cfg, H, W, prec):
    from oracle import model as om
    bp = util.blueprint(cfg)
    sd = util.cpu_state_dict(bp)
    imgs = torch.stack([util.make_image(i, H, W) for i in range(2)])
    out = bp.net(imgs.cuda())
    ref = om.forward(sd, util.oracle_cfg(cfg), imgs.float())
    S = out.S
    for s in range(len(ref.S)):
        assert S[s].shape == ref.S[s].shape
        assert _flip_stats(S[s].cpu(), ref.S[s]) < TOL[prec]['flips'], s
    assert torch.equal(S[0].cpu(), imgs.long())
    same = all(torch.equal(S[s].cpu(), ref.S[s]) for s in range(len(ref.S)))
    if same:      # no quantiser flips: parameters must agree to float noise
        for s in range(len(ref.P)):
            np.testing.assert_allclose(out.P[s].cpu().numpy(), ref.P[s].numpy(), rtol=TOL[prec]['p'],
                                       atol=TOL[prec]['p'])
    if cfg == 'cr':
        for s in range(1, 4):
            lev = sd['nets.0.enc.levels']
            assert torch.equal(out.bn[s].cpu(), lev[S[s].cpu()])


def test_get_P_reproduces_forward_bit_exactly(prec):
    """decoder side (three incremental get_P calls, other batch size) == encoder side."""
    bp = util.blueprint('cr')
    imgs = torch.stack([util.make_image(i, 48, 80) for i in range(3)]).cuda()
    out = bp.net(imgs)
    F_prev = None
    for s in (2, 1, 0):
        l, F_prev = bp.net.get_P_nhwc(s, out.bn8[s + 1][1:2].contiguous(), F_prev)
        assert torch.equal(l, out.P_nhwc[s][1:2]), s
    # reference-shaped entry point
    l, F = bp.net.get_P(2, out.bn[3][1:2])
    assert torch.equal(l, out.P[2][1:2])


def test_theoretical_bpsp_matches_golden(prec):
    g = util.golden_npz('l3c_32x32_i0')
    bp = util.blueprint('cr')
    img = torch.from_numpy(g['img']).unsqueeze(0).cuda()
    out = bp.forward(img)
    loss = bp.get_loss(out)
    np.testing.assert_allclose(loss.nonrecursive_bpsps, g['theory_bpsps'], rtol=2e-4 if prec in ('fp32', 'f16x2') else 2e-3)
    # per-sub-pixel map through the reference-shaped forward()
    dm = bp.losses.loss_dmol_rgb
    m = dm(img.float(), out.P[0])
    assert m.shape == (1, 3, 32, 32)
    np.testing.assert_allclose(float(m.sum()) / (np.log(2) * 3072), g['theory_bpsps'][0],
                               rtol=2e-4 if prec in ('fp32', 'f16x2') else 2e-3)


@pytest.mark.parametrize('name,cfg', [('l3c_32x32_i0', 'cr'), ('l3c_40x28_i1', 'cr'), ('rgbs_64x64_i0', 'cr_rgb_shared')])
def test_goldens_of_the_reference(name, cfg, tmp_path, prec):
    """(1) our decoder decodes the file the UNMODIFIED reference wrote, bit-exactly (needs the same
    symbols => same CDF integers along the coded path); (2) our encoder's file has the reference's
    size (bpsp within 1e-4 means: same byte count at this size) and layout; (3) round trip."""
    from l3c_pytorch_b200 import Bitcoding
    from l3c_pytorch_b200.codec import parse_container
    g = util.golden_npz(name)
    summ = util.golden_summary()[name]
    bp = util.blueprint(cfg)
    bc = Bitcoding(bp)
    img = torch.from_numpy(g['img'])
    p = str(tmp_path / 'ours.l3c')
    bpsp = bc.encode(img.long(), p)
    data = open(p, 'rb').read()
    dec = bc.decode(p)
    assert dec.dtype == torch.int64 and torch.equal(dec[0].cpu(), img.long())       # lossless
    ref = g['container'].tobytes()
    pt_o, sc_o = parse_container(data)
    pt_r, sc_r = parse_container(ref)
    assert pt_o == pt_r == tuple(summ['pad'] if 'pad' in summ else pt_r)
    assert [(C, H, W) for (C, H, W, _) in sc_o] == [(C, H, W) for (C, H, W, _) in sc_r]
    padded = 3 * (img.shape[1] + pt_o[2] + pt_o[3]) * (img.shape[2] + pt_o[0] + pt_o[1])
    assert abs(bpsp - len(data) * 8 / padded) < 1e-12
    # bpsp parity with the reference's own torchac path
    # (at these tiny sizes one byte is > 1e-4 bpsp: 2 bytes of slack for +-1-count CDF roundings; the
    # tf32 mode is only claimed at the 512^2 benchmark size, here it must merely stay close)
    slack = 2 + TOL[prec]['bytes_rel'] * summ['ref_bytes']
    assert abs(len(data) - summ['ref_bytes']) <= slack, (len(data), summ['ref_bytes'])


@pytest.mark.parametrize('name,cfg', [('l3c_32x32_i0', 'cr'), ('l3c_40x28_i1', 'cr'), ('rgbs_64x64_i0', 'cr_rgb_shared')])
def test_cross_decode_of_reference_files(name, cfg, tmp_path):
    """Decode the file the UNMODIFIED reference wrote (fp32 mode).  The uniform-prior scale is pure integer
    arithmetic and MUST decode to the reference's symbols.  The DMLL scales need bit-identical CDF
    integers at BOTH bounds of every coded symbol; CUDA expf/conv summation vs Sleef/MKL-DNN differ by one
    count in < 0.5 % of the entries (test_dmll_tables_and_intervals_vs_oracle), and the first differing
    bound desynchronises the coder for good -- the same holds between the reference's own CPU and GPU
    backends.  So the full cross-decode is asserted where it holds and xfail-ed WITH the measured first
    mismatch where it does not."""
    from l3c_pytorch_b200 import Bitcoding, engine as E
    g = util.golden_npz(name)
    bp = util.blueprint(cfg)
    bc = Bitcoding(bp)
    img = torch.from_numpy(g['img'])
    ref = g['container'].tobytes()
    old = E.get_conv_precision()
    E.set_conv_precision('fp32')
    try:
        # uniform scale through the codec's own path: decode everything, then look at the coarsest symbols
        from l3c_pytorch_b200.codec import parse_container
        _, scales = parse_container(ref)
        C, H, W, streams = scales[0]
        L = 256 if cfg != 'cr' else 25
        row = E.uniform_cdf_row(L)
        from l3c_pytorch_b200 import torchac
        cdf = torch.from_numpy(np.tile(row.view(np.int16), (1, H, W, 1)).copy())
        key = 'S3' if cfg == 'cr' else 'S1'
        for c, (off, n) in enumerate(streams):
            sym = torchac.decode_cdf(cdf, ref[off:off + n])
            assert torch.equal(sym.reshape(H, W).long(), torch.from_numpy(g[key][0, c].astype(np.int64))), c
        pr = str(tmp_path / 'ref.l3c')
        open(pr, 'wb').write(ref)
        dec_r = bc.decode(pr)[0].cpu()
    finally:
        E.set_conv_precision(old)
    if not torch.equal(dec_r, img.long()):
        bad = (dec_r != img.long())
        first = int(bad.flatten().nonzero()[0])
        pytest.xfail('cross-decode of the reference file desynchronised: %d of %d sub-pixels differ, first at '
                     'flat index %d (CDF integers differ by one count in <0.5%% of entries)'
                     % (int(bad.sum()), bad.numel(), first))


def test_round_trip_batch_512(prec):
    """BASELINE config 2, the WHOLE batch (16 x 3x512x512, image seeds 1000..1015): lossless, and bpsp
    of EVERY image within 1e-4 of what the unmodified reference writes for it (north-star tolerance;
    goldens: tests/golden/batch_bytes.json, made by oracle/gen_golden_batch.py) -- mean and max."""
    from l3c_pytorch_b200 import Bitcoding
    bp = util.blueprint('cr')
    bc = Bitcoding(bp)
    imgs = torch.stack([util.make_image(i, 512, 512) for i in range(16)])
    datas, bpsps = bc.encode_batch(imgs)
    dec = bc.decode_batch(datas)
    for i in range(16):
        assert torch.equal(dec[i][0].cpu(), imgs[i].long())
    gold = util.golden_batch_bytes()['l3c_512x512']
    assert [g['img_seed'] for g in gold] == list(range(1000, 1016))
    d = util.bpsp_deltas([len(x) for x in datas], [g['ref_bytes'] for g in gold], 3 * 512 * 512)
    print('512^2 batch parity [%s]: mean %+.2e, mean|.| %.2e, max|.| %.2e bpsp; bytes ours-ref %s'
          % (prec, d.mean(), np.abs(d).mean(), np.abs(d).max(),
             [len(x) - g['ref_bytes'] for x, g in zip(datas, gold)]))
    assert np.abs(d).max() <= 1e-4, d
    assert abs(bpsps[0] - gold[0]['ref_bpsp']) < 1e-4
    # batch-size independence of the bytes
    d1, _ = bc.encode_batch(imgs[:1])
    assert d1[0] == datas[0]


def test_crops_and_parts(tmp_path, monkeypatch):
    from l3c_pytorch_b200 import Bitcoding
    monkeypatch.setenv('AC_NEEDS_CROP_DIM', '40,40')
    bp = util.blueprint('cr')
    bc = Bitcoding(bp)
    g = torch.Generator().manual_seed(1000)
    img = (torch.rand(3, 100, 60, generator=g) * 255).round().to(torch.uint8)
    p = str(tmp_path / 'crop.l3c')
    bpsp = bc.encode(img.long(), p)
    parts = sorted(os.listdir(tmp_path))
    assert parts == ['crop.l3c.part%d' % i for i in range(4)]
    summ = util.golden_summary()['l3c_crop_100x60']
    sizes = [os.path.getsize(str(tmp_path / q)) for q in parts]
    assert open(str(tmp_path / parts[0]), 'rb').read()[:13].hex() == summ['header_hex']
    assert max(abs(a - b) for a, b in zip(sizes, summ['part_bytes'])) <= 3, (sizes, summ['part_bytes'])
    assert abs(bpsp - summ['ref_bpsp']) < 2e-3
    dec = bc.decode(str(tmp_path / parts[2]))          # any part name decodes + stitches everything
    assert torch.equal(dec[0].cpu(), img.long())


def test_rgb_shared_256(prec):
    """BASELINE config 4, the whole batch (32 x 3x256x256, cr_rgb_shared.cf) against the unmodified
    reference's container size of EVERY image.  One byte is 4.07e-5 bpsp here, and at random init nearly
    every symbol sits on a 1- or 2-count CDF interval, where a last-bit difference in a parameter (other
    summation order than MKL-DNN, CUDA vs Sleef expf) halves or doubles its probability: single images
    scatter by a few bytes either way (measured on the B200, fp32: -12 ... +5 bytes, mean |.| 3.4 bytes =
    1.4e-4 bpsp, i.e. SINGLE images are NOT within 1e-4 here; the batch mean is: -2.3e-5).  What must hold:
    no bias -- the batch MEAN within 1e-4 (fp32: the strict mode of this config) -- and no image further
    than 16 bytes (6.5e-4) from the reference.  The tf32 tensor-core mode is not claimed for this one-scale
    baseline (measured ~5e-4)."""
    from l3c_pytorch_b200 import Bitcoding
    bp = util.blueprint('cr_rgb_shared')
    bc = Bitcoding(bp)
    imgs = torch.stack([util.make_image(i, 256, 256) for i in range(32)])
    datas, bpsps = bc.encode_batch(imgs)
    dec = bc.decode_batch(datas)
    for i in range(32):
        assert torch.equal(dec[i][0].cpu(), imgs[i].long())
    gold = util.golden_batch_bytes()['rgbs_256x256']
    d = util.bpsp_deltas([len(x) for x in datas], [g['ref_bytes'] for g in gold], 3 * 256 * 256)
    print('rgb-shared 256^2 batch parity [%s]: mean %+.2e, mean|.| %.2e, max|.| %.2e bpsp; bytes ours-ref %s'
          % (prec, d.mean(), np.abs(d).mean(), np.abs(d).max(),
             [len(x) - g['ref_bytes'] for x, g in zip(datas, gold)]))
    if prec in ('fp32', 'f16x2'):              # f16x2: the tensor-core mode that IS claimed for this baseline
        assert abs(d.mean()) <= 1e-4, d
        assert np.abs(d).max() <= 16 * 8 / (3 * 256 * 256) + 1e-12, d
    else:
        assert abs(d.mean()) < 2e-3, d


def test_corrupt_file_is_detected(tmp_path):
    from l3c_pytorch_b200 import Bitcoding
    bp = util.blueprint('cr')
    bc = Bitcoding(bp)
    datas, _ = bc.encode_batch(util.make_image(3, 32, 32).unsqueeze(0))
    bad = bytearray(datas[0])
    bad[-1] ^= 0xFF
    with pytest.raises(ValueError):
        bc.decode_batch([bytes(bad)])


def test_reference_shaped_per_channel_api():
    """ArithmeticCoder + CodingCDFNonshared + cdf_step_non_shared, as the reference's
    code_with_cdf loop uses them (bitcoding.py:268-294), agree with the batched path."""
    from l3c_pytorch_b200 import ArithmeticCoder
    from l3c_pytorch_b200.coders import CodingCDFNonshared
    from l3c_pytorch_b200.codec import parse_container, BatchCodec
    bp = util.blueprint('cr')
    img = util.make_image(5, 32, 32).unsqueeze(0).cuda()
    out = bp.net(img)
    datas, info = BatchCodec(bp).encode_batch(img)
    _, scales = parse_container(datas[0])
    # RGB scale, channel by channel, through the reference-shaped calls
    dm = bp.losses.loss_dmol_rgb
    coding = CodingCDFNonshared(out.P[0], 3, dm)
    r = ArithmeticCoder(dm.L)
    dec_bn = torch.zeros(1, 3, 32, 32, device='cuda')
    for c in range(3):
        cdf = coding.get_next_C(dec_bn)
        enc = r.range_encode(out.S[0][:, c].to(torch.int16), cdf)
        off, n = scales[3][3][c]
        assert enc == datas[0][off:off + n], c
        back = r.range_decode(enc, cdf)
        assert torch.equal(back.long(), out.S[0][:, c].cpu())
        dec_bn[:, c] = img[:, c].float()


def test_cli_png_round_trip_with_checkpoint(tmp_path):
    """`l3c.py enc/dec` equivalent: checkpoint in the reference's {'net': state_dict} format, PNG in,
    .l3c out, PNG back -- pixel exact."""
    from PIL import Image
    from l3c_pytorch_b200 import cli
    bp = util.blueprint('cr')
    sd = {k: v.detach().cpu() + 0.001 for k, v in bp.net.state_dict().items()}     # not the default init
    ck = str(tmp_path / 'ckpt_0000000001.pt')
    torch.save({'net': sd}, ck)
    arr = util.make_image(7, 70, 50).permute(1, 2, 0).numpy()
    src = str(tmp_path / 'in.png')
    Image.fromarray(arr).save(src)
    out = str(tmp_path / 'x.l3c')
    assert cli.main(['--config', 'cr', '--ckpt', ck, 'enc', src, out]) == 0
    assert cli.main(['--config', 'cr', '--ckpt', ck, 'enc', src, out]) == 1          # exists -> EncodeError
    back = str(tmp_path / 'back.png')
    assert cli.main(['--config', 'cr', '--ckpt', ck, 'dec', out, back]) == 0
    assert (np.array(Image.open(back)) == arr).all()
    assert cli.main(['--config', 'cr', '--ckpt', ck, 'dec', str(tmp_path / 'missing.l3c'), back]) == 1


def test_tf32_and_fp32_modes_agree_on_size():
    """the tensor-core modes (tf32, f16) change parameters only by float noise: container sizes of a batch
    differ by a few bytes per image from the fp32 path (mean |d bpsp| < 1e-4 at 256^2 for L3C)."""
    from l3c_pytorch_b200 import Bitcoding, engine as E
    bp = util.blueprint('cr')
    bc = Bitcoding(bp)
    imgs = torch.stack([util.make_image(20 + i, 256, 256) for i in range(4)])
    old = E.get_conv_precision()
    try:
        sizes = {}
        for mode in ('fp32', 'tf32', 'f16'):
            E.set_conv_precision(mode)
            datas, _ = bc.encode_batch(imgs)
            dec = bc.decode_batch(datas)
            assert all(torch.equal(dec[i][0].cpu(), imgs[i].long()) for i in range(4))
            sizes[mode] = np.array([len(d) for d in datas])
    finally:
        E.set_conv_precision(old)
    for mode in ('tf32', 'f16'):
        d_bpsp = np.abs(sizes['fp32'] - sizes[mode]).mean() * 8 / (3 * 256 * 256)
        assert d_bpsp < 1e-4, (mode, sizes, d_bpsp)


def test_full_size_crop_config_round_trip(tmp_path):
    """BASELINE config 5 shape: one 3x3000x2000 image -> 4 crops of 1500x1000 padded to 1504x1000,
    coded as one batch into .part0..3, decoded and stitched -- lossless (size-independent property)."""
    from l3c_pytorch_b200 import Bitcoding, engine as E
    bp = util.blueprint('cr')
    bc = Bitcoding(bp)
    old = E.get_conv_precision()
    E.set_conv_precision('f16')
    try:
        g = torch.Generator().manual_seed(4242)
        img = (torch.rand(3, 3000, 2000, generator=g) * 255).round().to(torch.uint8)
        p = str(tmp_path / 'big.l3c')
        bpsp = bc.encode(img.long(), p)
        parts = sorted(os.listdir(tmp_path))
        assert parts == ['big.l3c.part%d' % i for i in range(4)]
        head = open(str(tmp_path / parts[0]), 'rb').read()[:8]
        assert head == bytes([0, 0, 0, 0, 2, 0, 2, 0])                  # pad (L,R,T,B) = (0,0,2,2): 1504x1000
        assert 17.5 < bpsp < 19.0
        dec = bc.decode(str(tmp_path / parts[0]))
        assert dec.shape == (1, 3, 3000, 2000) and torch.equal(dec[0].cpu(), img.long())
    finally:
        E.set_conv_precision(old)


def test_rgb_shared_batch32_round_trip():
    """BASELINE config 4 shape: 32 x 3x256x256, cr_rgb_shared.cf (strict fp32 mode)."""
    from l3c_pytorch_b200 import Bitcoding, engine as E
    bp = util.blueprint('cr_rgb_shared')
    bc = Bitcoding(bp)
    old = E.get_conv_precision()
    E.set_conv_precision('fp32')
    try:
        imgs = torch.stack([util.make_image(100 + i, 256, 256) for i in range(32)])
        datas, bpsps = bc.encode_batch(imgs)
        dec = bc.decode_batch(datas)
        assert all(torch.equal(dec[i][0].cpu(), imgs[i].long()) for i in range(32))
        assert abs(np.mean(bpsps) - 15.25) < 0.05
    finally:
        E.set_conv_precision(old)


def test_async_encode_beside_a_decode_gives_the_same_bytes():
    """encode_batch_begin on the side stream (its own SM partition where the driver offers one) while
    another batch is being decoded: same containers as the synchronous call, lossless both ways."""
    import l3c_pytorch_b200 as l3c
    bp = util.blueprint('cr')
    bc = l3c.Bitcoding(bp)
    a = torch.stack([util.make_image(i, 128, 128) for i in range(4)])
    b = torch.stack([util.make_image(10 + i, 128, 128) for i in range(4)])
    datas_a, _ = bc.encode_batch(a)
    want_b, bpsp_b = bc.encode_batch(b)
    side = bc.side_stream(a.shape[0])
    job = bc.encode_batch_begin(b.pin_memory(), stream=side)         # enqueued, not waited for
    dec_a = bc.decode_batch(datas_a)                                  # runs beside it
    got_b, bpsp_b2 = job.finish()
    assert got_b == want_b and bpsp_b2 == bpsp_b
    for i in range(4):
        assert torch.equal(dec_a[i][0].cpu(), a[i].long())
    dec_b = bc.decode_batch(got_b)
    for i in range(4):
        assert torch.equal(dec_b[i][0].cpu(), b[i].long())
    # several jobs in flight on the same stream finish in any order
    jobs = [bc.encode_batch_begin(x, stream=side) for x in (a, b)]
    assert jobs[1].finish()[0] == want_b and jobs[0].finish()[0] == datas_a


def test_decodes_in_flight_on_two_lanes():
    """Several decode_batch calls in flight, each on its own DecodeLane (the range decoders of all of them on
    one SM partition), beside an encode on the side stream: every image comes back bit-exact and the
    containers equal the synchronous ones."""
    import l3c_pytorch_b200 as l3c
    bp = util.blueprint('cr')
    bc = l3c.Bitcoding(bp)
    sets = [torch.stack([util.make_image(40 * k + i, 96, 160) for i in range(3)]) for k in range(4)]
    datas = [bc.encode_batch(x)[0] for x in sets]
    lanes = bc.decode_lanes(3, n_lanes=2)
    assert len(lanes) == 2 and lanes[0].main.cuda_stream != lanes[1].main.cuda_stream
    side = bc.side_stream(3, n_lanes=2)
    cur = torch.cuda.current_stream()
    job = bc.encode_batch_begin(sets[0].pin_memory(), stream=side)
    outs = []
    for k in range(4):                                        # four decodes queued on two lanes, none awaited
        ln = lanes[k % 2]
        ln.main.wait_stream(cur)
        with torch.cuda.stream(ln.main):
            outs.append(bc.decode_batch(datas[k], lane=ln))
    assert job.finish()[0] == datas[0]
    for ln in lanes:
        cur.wait_stream(ln.main)
    torch.cuda.synchronize()
    for k in range(4):
        for i in range(3):
            assert torch.equal(outs[k][i][0].cpu(), sets[k][i].long()), (k, i)


def test_sm_partition_streams_run_kernels_and_report_group_sizes():
    """l3c_partition_streams: two disjoint SM groups; a kernel launched into either gives the same
    result as on an ordinary stream.  Skipped where the driver has no green contexts."""
    from l3c_pytorch_b200 import engine as E
    dev = torch.device('cuda', torch.cuda.current_device())
    part = E.partition_streams(dev, 24, 3, 4)
    if part is None:
        pytest.skip('driver cannot partition SMs (L3C_EUNSUPPORTED)')
    s_a, s_b, n_a, n_b = part
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    assert len(s_a) == 3 and len(s_b) == 4 and n_a >= 24 and n_b >= 8 and n_a + n_b <= n_sm
    assert E.partition_streams(dev, 24, 3, 4)[0][0].cuda_stream == s_a[0].cuda_stream     # cached, not re-created
    bp = util.blueprint('cr')
    conv = bp.net.nets[0].enc.body[0].body[0]
    x = torch.randn(2, 64, 64, 64, device=dev)
    old = E.get_conv_precision()
    E.set_conv_precision('tf32')
    try:
        want = E.conv2d(conv, x)
        torch.cuda.synchronize()
        for st in (s_a[0], s_b[0], s_b[3]):
            with torch.cuda.stream(st):
                got = E.conv2d(conv, x)                      # persistent grid sized to the group
            st.synchronize()
            assert torch.equal(got, want)
    finally:
        E.set_conv_precision(old)


def test_crafted_container_headers_raise_before_the_kernels_run():
    """C/H/W come from the file; a wrong-config or crafted header must raise ValueError (the
    reference fails with a torch shape error) -- never reach a kernel with buffers sized from the file."""
    import struct
    from l3c_pytorch_b200 import Bitcoding
    from l3c_pytorch_b200.codec import parse_container
    bp = util.blueprint('cr')
    bc = Bitcoding(bp)
    datas, _ = bc.encode_batch(util.make_image(3, 32, 32).unsqueeze(0))
    good = datas[0]
    _, scales = parse_container(good)
    # header of the finest scale: u8 C, u16 H, u16 W sits 5 bytes before its first stream-length field
    off_rgb = scales[3][3][0][0] - 4 - 5
    assert struct.unpack_from('<BHH', good, off_rgb) == (3, 32, 32)
    for patch in [(3, 16, 16), (3, 32, 64), (3, 64, 64)]:
        bad = bytearray(good)
        struct.pack_into('<BHH', bad, off_rgb, *patch)
        with pytest.raises(ValueError):
            bc.decode_batch([bytes(bad)])
    off_z1 = scales[2][3][0][0] - 4 - 5
    assert struct.unpack_from('<BHH', good, off_z1) == (5, 16, 16)
    bad = bytearray(good)
    struct.pack_into('<BHH', bad, off_z1, 5, 8, 8)
    with pytest.raises(ValueError):
        bc.decode_batch([bytes(bad)])
    # a file of the RGB-shared model given to the L3C model: wrong number of scales
    bcs = Bitcoding(util.blueprint('cr_rgb_shared'))
    ds, _ = bcs.encode_batch(util.make_image(3, 32, 32).unsqueeze(0))
    with pytest.raises(ValueError):
        bc.decode_batch(ds)
    with pytest.raises(ValueError):
        bcs.decode_batch([good])
    assert torch.equal(bc.decode_batch([good])[0][0].cpu(), util.make_image(3, 32, 32).long())


def test_part_files_are_not_overwritten(tmp_path, monkeypatch):
    from l3c_pytorch_b200 import Bitcoding
    monkeypatch.setenv('AC_NEEDS_CROP_DIM', '40,40')
    bc = Bitcoding(util.blueprint('cr'))
    img = util.make_image(0, 100, 60)
    p = str(tmp_path / 'crop.l3c')
    bc.encode(img.long(), p)
    with pytest.raises(AssertionError):
        bc.encode(img.long(), p)                                      # part0..3 exist
    q = str(tmp_path / 'other.l3c')
    open(q + '.part7', 'wb').close()                                  # a stale higher-numbered part
    with pytest.raises(AssertionError):
        bc.encode(img.long(), q)


def test_recursive_theoretical_bpsp_matches_reference():
    """`--recursive` evaluation of the RGB-shared baseline (the shared scale applied
    again to its own thumbnails, multiscale_network.py:235-238,291-294) against per-scale costs of the
    UNMODIFIED reference (tests/golden/recursive.json, oracle/gen_golden_recursive.py)."""
    import json
    from l3c_pytorch_b200 import engine as E
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'recursive.json')) as f:
        cases = json.load(f)['cases']
    bp = util.blueprint('cr_rgb_shared')
    old = E.get_conv_precision()
    E.set_conv_precision('fp32')
    try:
        for c in cases:
            img = util.make_image(c['img_index'], c['H'], c['W']).unsqueeze(0).cuda()
            out = bp.forward(img, auto_recurse=c['auto_recurse'])
            assert [list(s.shape) for s in out.S] == c['S_shapes']
            assert out.auto_recursive_from == 1
            loss = bp.get_loss(out)
            np.testing.assert_allclose(loss.nonrecursive_bpsps, c['nonrecursive_bpsps'], rtol=3e-4)
            np.testing.assert_allclose(loss.recursive_bpsps, c['recursive_bpsps'], rtol=3e-4)
            per_img = bp.get_loss_per_image(out)
            assert abs(per_img[0] - sum(c['recursive_bpsps'])) < 3e-4 * sum(c['recursive_bpsps'])
        # without recursion nothing changes
        out0 = bp.forward(img)
        assert out0.auto_recursive_from is None and bp.get_loss(out0).recursive_bpsps is None
        # L3C proper: the flag is ignored by the tester (multiscale_tester.py:123-125) but the network supports it
        bpl = util.blueprint('cr')
        o = bpl.forward(util.make_image(0, 64, 64).unsqueeze(0).cuda(), auto_recurse=1)
        assert len(o.S) == 5 and o.auto_recursive_from == 3
    finally:
        E.set_conv_precision(old)


def test_sampling_kernel_and_sample_forward():
    """`dmll_sample_kernel` against a plain PyTorch restatement of
    logistic_mixture.py:277-323 fed with the SAME uniform random numbers, then the whole sample_forward
    (multiscale_network.py:328-406) for the three styles test.py --sample writes."""
    from l3c_pytorch_b200.dmll import DiscretizedMixLogisticLoss
    g = torch.Generator().manual_seed(7)
    for rgb, C, x_min, x_max, L in ((True, 3, 0, 255, 256), (False, 5, -1, 1, 25)):
        K, N, H, W = 10, 2, 9, 13
        P = 4 if rgb else 3
        l = torch.randn(N, P * C * K, H, W, generator=g) * 1.5
        lr = l.reshape(N, P, C, K, H, W)
        if rgb:
            lr[:, 1] = lr[:, 1] * 40 + 128
            lr[:, 2] = lr[:, 2] + 1.5
        u_sel = torch.empty(N, C, K, H, W).uniform_(1e-5, 1 - 1e-5, generator=g)
        u_x = torch.empty(N, C, H, W).uniform_(1e-5, 1 - 1e-5, generator=g)
        dm = DiscretizedMixLogisticLoss(rgb, x_min, x_max, L)
        got = dm.sample(l.cuda(), C, u_sel.cuda(), u_x.cuda()).cpu()
        sel = torch.argmax(lr[:, 0] - torch.log(-torch.log(u_sel)), dim=2).unsqueeze(2)
        means = torch.gather(lr[:, 1], 2, sel).squeeze(2)
        ls = torch.clamp(torch.gather(lr[:, 2], 2, sel).squeeze(2), min=-7.)
        x = means + torch.exp(ls) * (torch.log(u_x) - torch.log(1. - u_x))
        if rgb:
            co = torch.sigmoid(lr[:, 3])
            sg, sb = sel[:, 1], sel[:, 2]
            c_gr = torch.gather(co[:, 0], 1, sg).squeeze(1)
            c_br = torch.gather(co[:, 1], 1, sb).squeeze(1)
            c_bg = torch.gather(co[:, 2], 1, sb).squeeze(1)
            x0 = x[:, 0].clamp(0, 255.)
            x1 = (x[:, 1] + c_gr * x0).clamp(0, 255.)
            x2 = (x[:, 2] + c_br * x0 + c_bg * x1).clamp(0, 255.)
            x = torch.stack((x0, x1, x2), dim=1)
        assert got.shape == x.shape
        np.testing.assert_allclose(got.numpy(), x.numpy(), rtol=2e-4, atol=2e-3)
    bp = util.blueprint('cr')
    img = util.make_image(0, 64, 96).unsqueeze(0).cuda()
    for scales in ([], [0], [0, 1], [0, 1, 2]):
        s = bp.sample_forward(img, scales)
        assert s.shape == (1, 3, 64, 96) and s.dtype == torch.float32
        assert float(s.min()) >= 0.0 and float(s.max()) <= 255.0 and bool(torch.isfinite(s).all())
    s2 = bp.sample_forward(img, [0, 1, 2], partial_final=[0, 1])
    assert s2.shape == (1, 3, 64, 96)


@pytest.mark.parametrize('cfg,H,W,tile', [('cr', 64, 96, (16, 16)), ('cr', 128, 128, (64, 64)), ('cr', 72, 104, (32, 48)),
                                          ('cr_rgb_shared', 64, 64, (16, 32))])
def test_tiled_containers_round_trip(cfg, H, W, tile):
    """Throughput mode (compat AND tiled): every channel plane of every scale cut into tiles
    that are coded as independent streams.  Lossless; decode recognises the layout from the container; the
    symbols coded are the same, so the size differs from the reference layout only by the per-stream overhead
    (4-byte length + termination: <= 8 bytes per extra stream)."""
    from l3c_pytorch_b200 import Bitcoding, engine as E
    from l3c_pytorch_b200.codec import parse_container, container_tile, tile_grid
    bp = util.blueprint(cfg)
    old = E.get_conv_precision()
    E.set_conv_precision('f16' if cfg == 'cr' else 'fp32')
    try:
        imgs = torch.stack([util.make_image(i, H, W) for i in range(3)])
        compat, _ = Bitcoding(bp).encode_batch(imgs)
        bct = Bitcoding(bp, tile=tile)
        tiled, bpsps = bct.encode_batch(imgs)
        assert container_tile(tiled[0]) == tile and container_tile(compat[0]) is None
        dec = Bitcoding(bp).decode_batch(tiled)                 # the layout comes from the container, not the object
        for i in range(3):
            assert torch.equal(dec[i][0].cpu(), imgs[i].long())
        _, sc = parse_container(tiled[0])
        _, sc0 = parse_container(compat[0])
        n_streams = sum(len(st) for (_, _, _, st) in sc)
        assert n_streams == sum(C * len(tile_grid(h, w, tile)) for (C, h, w, _) in sc0) > sum(len(st) for (_, _, _, st) in sc0)
        extra = n_streams - sum(len(st) for (_, _, _, st) in sc0)
        for a, b in zip(compat, tiled):
            assert 0 <= len(b) - len(a) - 8 <= 8 * extra, (len(a), len(b), extra)
        # mixed batch of both layouts through the file API
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            p = os.path.join(tmp, 'x.l3c')
            bct.encode(imgs[0].long(), p)
            assert torch.equal(Bitcoding(bp).decode(p)[0].cpu(), imgs[0].long())
    finally:
        E.set_conv_precision(old)
