"""Single-image command line: the encode/decode half of the reference's `l3c.py`
(/root/reference/src/l3c.py:74-129) on the B200 path.

    python -m l3c_pytorch_b200.cli --config cr --ckpt CKPT.pt enc IMG OUT.l3c [--overwrite]
    python -m l3c_pytorch_b200.cli --config cr --ckpt CKPT.pt dec IN.l3c OUT.png

`--ckpt` is a checkpoint as written by the reference's Saver (`torch.save({'net': state_dict})`,
helpers/saver.py:168); the module tree here has the same parameter names, so it loads strict.
Without `--ckpt` the seed-0 default initialisation is used (only useful for testing).
Images larger than 2000x1500 are coded as crops into `OUT.l3c.part{i}` exactly as the reference does.
"""
import argparse
import os
import sys

import numpy as np
import torch


class EncodeError(Exception):
    pass


class DecodeError(Exception):
    pass


def load_blueprint(config_name, ckpt=None, precision=None, device='cuda:0'):
    from . import config, engine
    from .blueprint import MultiscaleBlueprint
    if precision:
        engine.set_conv_precision(precision)
    torch.manual_seed(0)
    bp = MultiscaleBlueprint(config.ms_config(config_name), device=device)
    bp.set_eval()
    if ckpt:
        state = torch.load(ckpt, map_location='cpu')
        bp.net.load_state_dict(state['net'] if 'net' in state else state, strict=True)   # saver.py:184-210
        bp.net.to(bp.device)
    return bp


def read_image(path):
    """PIL -> int64 1CHW, alpha dropped (multiscale_tester.py:411-422)."""
    from PIL import Image
    img = np.array(Image.open(path))
    if img.ndim == 2:
        img = np.stack([img] * 3, -1)
    img = img[..., :3]
    return torch.from_numpy(img).permute(2, 0, 1).unsqueeze(0).long()


def write_image(img_1chw, path):
    from PIL import Image
    assert path.endswith('.png'), 'decoded images are written as PNG (lossless)'
    arr = img_1chw[0].permute(1, 2, 0).to(torch.uint8).cpu().numpy()
    Image.fromarray(arr).save(path)


def encode(bc, img_p, pout, overwrite=False):
    if os.path.isfile(pout) or os.path.isfile(pout + '.part0'):
        if not overwrite:
            raise EncodeError('{} exists. Consider --overwrite.'.format(pout))
        for q in [pout] + [pout + '.part%d' % i for i in range(4096)]:
            if os.path.isfile(q):
                os.remove(q)
            elif q != pout:
                break
    img = read_image(img_p)
    bpsp = bc.encode(img, pout)
    print('Encoded {} -> {}: {:.4f} bpsp'.format(img_p, pout, bpsp))
    return bpsp


def decode(bc, pin, png_out):
    if not os.path.isfile(pin):
        if os.path.isfile(pin + '.part0'):
            pin = pin + '.part0'
        else:
            raise DecodeError('File not found: {}'.format(pin))
    img = bc.decode(pin)
    write_image(img, png_out)
    print('Decoded {} -> {}'.format(pin, png_out))


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('--config', default='cr', help='cr | cr_rgb_shared | cr_rgb or a path to a .cf file')
    p.add_argument('--ckpt', default=None, help="checkpoint with {'net': state_dict}")
    p.add_argument('--precision', default=None, choices=['fp32', 'tf32', 'f16', 'f16x2'],
                   help='conv mode (default fp32; encoder and decoder must use the same one)')
    sub = p.add_subparsers(dest='mode', required=True)
    e = sub.add_parser('enc')
    e.add_argument('img')
    e.add_argument('out')
    e.add_argument('--overwrite', action='store_true')
    d = sub.add_parser('dec')
    d.add_argument('inp')
    d.add_argument('out_png')
    a = p.parse_args(argv)
    from .bitcoding import Bitcoding
    bc = Bitcoding(load_blueprint(a.config, a.ckpt, a.precision))
    try:
        if a.mode == 'enc':
            encode(bc, a.img, a.out, a.overwrite)
        else:
            decode(bc, a.inp, a.out_png)
    except (EncodeError, DecodeError) as err:
        print('*** ERROR: {}'.format(err))
        return 1
    return 0


if __name__ == '__main__':
    sys.exit(main())
