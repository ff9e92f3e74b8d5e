"""oracle/gen_golden_recursive.py -- TEST INFRASTRUCTURE ONLY.  Run in the build container (needs
/root/reference):   python oracle/gen_golden_recursive.py

Golden for SURVEY.md section 8 row f2: the theoretical bpsp of the RGB-shared baseline evaluated with
`--recursive` (auto_recurse = 3 more applications of the shared scale).  Imports the UNMODIFIED reference
(multiscale_network.py:226-306, multiscale_blueprint.py:64-95) on the CPU, seed-0 default-init weights,
and writes per-scale `nonrecursive_bpsps` / `recursive_bpsps` of a few seeded images to
tests/golden/recursive.json.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  (sets up sys.path for the reference and its shims)

import torch  # noqa: E402


def main():
    cfg, bp, bc, _ = gg.load_reference('cr_rgb_shared.cf')
    cases = []
    for (H, W, idx, rec) in [(64, 64, 0, 3), (96, 128, 1, 3), (64, 64, 2, 1), (128, 128, 3, 'padded')]:
        img = gg.make_image(idx, H, W)
        r = 3 if rec == 'padded' else rec
        out = bp.forward(img.unsqueeze(0).float(), auto_recurse=r)
        loss = bp.get_loss(out)
        cases.append({'H': H, 'W': W, 'img_index': idx, 'auto_recurse': r,
                      'nonrecursive_bpsps': [float(x) for x in loss.nonrecursive_bpsps],
                      'recursive_bpsps': [float(x) for x in loss.recursive_bpsps],
                      'S_shapes': [list(s.shape) for s in out.S]})
        print(cases[-1])
    p = os.path.join(gg.ROOT, 'tests', 'golden', 'recursive.json')
    with open(p, 'w') as f:
        json.dump({'config': 'cr_rgb_shared.cf', 'weights': 'torch.manual_seed(0) default init',
                   'cases': cases}, f, indent=1)
    print('wrote', p)


if __name__ == '__main__':
    main()
