"""oracle/gen_golden.py -- TEST INFRASTRUCTURE ONLY.  Run in the build container (needs
/root/reference):   python oracle/gen_golden.py

Imports the UNMODIFIED reference Python from /root/reference/src (with stand-ins for its
un-vendored deps, oracle/ref_shims) plus its own torchac.cpp compiled into oracle/_ref, runs the
reference's Bitcoding encode/decode on seeded inputs, cross-checks oracle/model.py against it, and
writes small fixtures into tests/golden/.  The two decode-side monkey-patches P1/P2 (SURVEY.md
section 8c) are applied; they do not change encoder output.
"""
import hashlib
import json
import os
import sys
import tempfile

os.environ.setdefault('PYTHONDONTWRITEBYTECODE', '1')
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = '/root/reference/src'
sys.path.insert(0, os.path.join(HERE, '_ref'))
sys.path.insert(0, os.path.join(HERE, 'ref_shims'))
sys.path.insert(0, REF_SRC)
sys.path.insert(1, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_grad_enabled(False)


def make_image(i, H, W):
    """BASELINE.md recipe: uniform-noise uint8 image i."""
    g = torch.Generator().manual_seed(1000 + i)
    return (torch.rand(3, H, W, generator=g) * 255).round().to(torch.uint8)


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def load_reference(cfg_name, crop_dim=None):
    if crop_dim:
        os.environ['AC_NEEDS_CROP_DIM'] = crop_dim
    from oracle import build_ref
    assert build_ref.build() is not None
    import pytorch_ext as pe
    pe.CUDA_AVAILABLE = False
    pe.set_device(False)
    from fjcommon import config_parser, no_op
    from blueprints.multiscale_blueprint import MultiscaleBlueprint
    from bitcoding.bitcoding import Bitcoding
    import bitcoding.bitcoding as bcmod
    from criterion import logistic_mixture as lm

    cfg, _ = config_parser.parse(os.path.join(REF_SRC, 'configs/ms', cfg_name))
    torch.manual_seed(0)
    bp = MultiscaleBlueprint(cfg)
    bp.set_eval()

    # P1: the bit-decoder must map symbols through the same linspace LUT as the quantizer
    def to_bn_p1(self, S):
        if self.L == 256:
            return S.float()
        return torch.linspace(self.x_min, self.x_max, self.L)[S.long()]
    lm.DiscretizedMixLogisticLoss.to_bn = to_bn_p1

    if cfg.rgb_bicubic_baseline:
        # P2: RGB baselines: decoder must feed S - rgb_mean (what the encoder fed)
        orig_decode_uniform = Bitcoding.decode_uniform
        mean = bp.net.nets[0].enc.rgb_mean

        class _Shift(object):
            def __init__(self, dm):
                self.dm = dm

            def __getattr__(self, k):
                return getattr(self.dm, k)

        def decode_patched(self, pin, _recurse_part=True):
            dm = self.blueprint.losses.loss_dmol_n
            old = dm.to_bn
            try:
                # only the uniform-scale to_bn sees whole S tensors with C==3 at the first call
                calls = {'n': 0}

                def tb(S):
                    calls['n'] += 1
                    if calls['n'] == 1:
                        return S.float() - mean
                    return S.float()
                dm.to_bn = tb
                return orig_decode(self, pin, _recurse_part)
            finally:
                dm.to_bn = old
        orig_decode = Bitcoding.decode
        Bitcoding.decode = decode_patched
        del orig_decode_uniform
    bc = Bitcoding(bp, times=no_op.NoOp)
    return cfg, bp, bc, bcmod


def run_case(bp, bc, img_u8, tmp, name):
    p = os.path.join(tmp, name + '.l3c')
    bpsp = bc.encode(img_u8.long(), p)
    parts = sorted([q for q in os.listdir(tmp) if q.startswith(name + '.l3c')],
                   key=lambda s: (len(s), s))
    datas = [open(os.path.join(tmp, q), 'rb').read() for q in parts]
    dec = bc.decode(os.path.join(tmp, parts[0]))
    assert (dec[0] == img_u8.long()).all(), 'reference round trip failed for ' + name
    return float(bpsp), datas


def main():
    from oracle import model as om
    gold = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(gold, exist_ok=True)
    summary = {}

    # ------------------------------------------------------------------ L3C (cr.cf)
    cfg, bp, bc, _ = load_reference('cr.cf')
    sd = {k: v.detach().clone() for k, v in bp.net.state_dict().items()}
    summary['l3c_sd_sha256'] = sd_digest(sd)
    summary['l3c_sd_numel'] = int(sum(v.numel() for v in sd.values()))
    summary['l3c_sd_probe'] = {k: [float(x) for x in sd[k].flatten()[:3]]
                               for k in ['heads.0.head.1.head.weight', 'nets.2.dec.tail.0.bias',
                                         'prob_clfs.0.atrous.lin.weight']}
    with tempfile.TemporaryDirectory() as tmp:
        for (H, W, idx) in [(32, 32, 0), (40, 28, 1), (128, 128, 0), (512, 512, 0)]:
            name = 'l3c_%dx%d_i%d' % (H, W, idx)
            img = make_image(idx, H, W)
            bpsp, datas = run_case(bp, bc, img, tmp, name)
            assert len(datas) == 1
            data = datas[0]
            # --- cross-check the oracle restatement against the reference
            d_t, dbg_t = om.encode_image(sd, om.CFG_L3C, img, 'torch', return_debug=True)
            assert d_t == data, 'oracle (torch CDF formula) != reference bytes for ' + name
            d_k, dbg_k = om.encode_image(sd, om.CFG_L3C, img, 'kernel', return_debug=True)
            if H <= 128:
                assert (om.decode_image(sd, om.CFG_L3C, d_k, 'kernel')[0] == img.long()).all()
            entry = dict(H=H, W=W, img_seed=1000 + idx, ref_bytes=len(data), ref_bpsp=bpsp,
                         ref_sha256=hashlib.sha256(data).hexdigest(),
                         oracle_kernel_formula_bytes=len(d_k),
                         oracle_kernel_formula_sha256=hashlib.sha256(d_k).hexdigest(),
                         stream_bytes_ref={'%d_%d' % k: len(v) for k, v in dbg_t['streams'].items()},
                         stream_bytes_kernel={'%d_%d' % k: len(v) for k, v in dbg_k['streams'].items()},
                         pad=list(dbg_t['pad']))
            # theoretical bpsp through the reference's own get_loss
            pt = dbg_t['pad']
            imgp = torch.nn.functional.pad(img.unsqueeze(0).long(), pt).float()
            out_ref = bp.forward(imgp)
            loss = bp.get_loss(out_ref)
            entry['ref_theory_bpsps'] = [float(x) for x in loss.nonrecursive_bpsps]
            out_or = dbg_t['out']
            out_ref = bp.forward(imgp)          # get_loss mutates nothing, but P[0] views are fresh
            for s in range(4):
                assert (out_ref.S[s] == out_or.S[s]).all()
            for s in range(3):
                assert torch.equal(out_ref.P[s], out_or.P[s]), 'oracle forward != reference'
            th = om.theoretical_bpsps(om.CFG_L3C, out_or)
            assert np.allclose(th, entry['ref_theory_bpsps'], rtol=1e-5), (th, entry['ref_theory_bpsps'])
            summary[name] = entry
            print(name, entry['ref_bytes'], entry['oracle_kernel_formula_bytes'], bpsp)
            if (H, W) in [(32, 32), (40, 28)]:
                rgbdm, odm = om.dmlls(om.CFG_L3C)
                zeros = torch.zeros_like(out_or.bn[1])
                np.savez_compressed(
                    os.path.join(gold, name + '.npz'),
                    img=img.numpy(), container=np.frombuffer(data, np.uint8),
                    container_kernel_formula=np.frombuffer(d_k, np.uint8),
                    S1=out_or.S[1].numpy().astype(np.int8), S2=out_or.S[2].numpy().astype(np.int8),
                    S3=out_or.S[3].numpy().astype(np.int8),
                    P0=out_or.P[0].numpy()[:, :, ::4, ::4], P1=out_or.P[1].numpy()[:, :, ::2, ::2],
                    P2=out_or.P[2].numpy(),
                    cdf_z1_c0=om.cdf_table_kernel_formula(odm, out_or.P[1], 0, 5, zeros)[::7],
                    cdf_z1_c0_torch=om.cdf_table_torch_formula(odm, out_or.P[1], 0, 5, zeros)[::7],
                    theory_bpsps=np.array(entry['ref_theory_bpsps']))

    # ------------------------------------------------------------------ crops (cr.cf)
    for m in [m for m in sys.modules if m.split('.')[0] in
              ('auto_crop', 'bitcoding', 'blueprints', 'modules', 'criterion', 'helpers', 'pytorch_ext',
               'vis', 'test', 'dataloaders')]:
        del sys.modules[m]
    cfg, bp, bc, _ = load_reference('cr.cf', crop_dim='40,40')
    assert sd_digest(bp.net.state_dict()) == summary['l3c_sd_sha256']
    with tempfile.TemporaryDirectory() as tmp:
        g = torch.Generator().manual_seed(1000)
        img = (torch.rand(3, 100, 60, generator=g) * 255).round().to(torch.uint8)
        bpsp, datas = run_case(bp, bc, img, tmp, 'crop')
        summary['l3c_crop_100x60'] = dict(ref_bpsp=bpsp, part_bytes=[len(d) for d in datas],
                                          part_sha256=[hashlib.sha256(d).hexdigest() for d in datas],
                                          header_hex=datas[0][:13].hex())
        print('crop', summary['l3c_crop_100x60'])
    os.environ.pop('AC_NEEDS_CROP_DIM')

    # ------------------------------------------------------------------ RGB shared
    for m in [m for m in sys.modules if m.split('.')[0] in
              ('auto_crop', 'bitcoding', 'blueprints', 'modules', 'criterion', 'helpers', 'pytorch_ext',
               'vis', 'test', 'dataloaders')]:
        del sys.modules[m]
    cfg, bp, bc, _ = load_reference('cr_rgb_shared.cf')
    sd = {k: v.detach().clone() for k, v in bp.net.state_dict().items()}
    summary['rgbs_sd_sha256'] = sd_digest(sd)
    summary['rgbs_sd_numel'] = int(sum(v.numel() for v in sd.values()))
    with tempfile.TemporaryDirectory() as tmp:
        for (H, W, idx) in [(64, 64, 0), (256, 256, 0)]:
            name = 'rgbs_%dx%d_i%d' % (H, W, idx)
            img = make_image(idx, H, W)
            bpsp, datas = run_case(bp, bc, img, tmp, name)
            data = datas[0]
            d_t, dbg_t = om.encode_image(sd, om.CFG_RGB_SHARED, img, 'torch', return_debug=True)
            assert d_t == data, 'oracle != reference bytes for ' + name
            d_k = om.encode_image(sd, om.CFG_RGB_SHARED, img, 'kernel')
            if H <= 64:
                assert (om.decode_image(sd, om.CFG_RGB_SHARED, d_k)[0] == img.long()).all()
            summary[name] = dict(H=H, W=W, ref_bytes=len(data), ref_bpsp=bpsp,
                                 ref_sha256=hashlib.sha256(data).hexdigest(),
                                 oracle_kernel_formula_bytes=len(d_k))
            print(name, summary[name])
            if H == 64:
                np.savez_compressed(os.path.join(gold, name + '.npz'), img=img.numpy(),
                                    container=np.frombuffer(data, np.uint8),
                                    S1=dbg_t['out'].S[1].numpy().astype(np.uint8),
                                    P0=dbg_t['out'].P[0].numpy()[:, :, ::4, ::4])

    with open(os.path.join(gold, 'summary.json'), 'w') as f:
        json.dump(summary, f, indent=1, sort_keys=True)
    print('wrote', gold)


if __name__ == '__main__':
    main()
