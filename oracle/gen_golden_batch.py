"""oracle/gen_golden_batch.py -- TEST INFRASTRUCTURE ONLY.  Run in the build container (needs
/root/reference):   python oracle/gen_golden_batch.py

Batch-wide bpsp goldens for BASELINE configs 2 and 4: the container size the UNMODIFIED reference
(`Bitcoding.encode`, /root/reference/src/bitcoding/bitcoding.py:50-123, with its own compiled
torchac.cpp) writes for EVERY image of the benchmark batches --

  * L3C `cr.cf`, 3x512x512, image seeds 1000..1015          (config 2 / 3: 16 images per GPU)
  * RGB-shared `cr_rgb_shared.cf`, 3x256x256, seeds 1000..1031 (config 4: 32 images)

written to tests/golden/batch_bytes.json.  Encoder output is not affected by the decode-side
patches P1/P2, so nothing is patched here: the reference runs as it is.
"""
import json
import os
import sys
import tempfile
import time

os.environ.setdefault('PYTHONDONTWRITEBYTECODE', '1')
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  (sets sys.path for the reference + shims)
import torch  # noqa: E402

torch.set_grad_enabled(False)


def _forget_reference():
    for m in [m for m in sys.modules if m.split('.')[0] in
              ('auto_crop', 'bitcoding', 'blueprints', 'modules', 'criterion', 'helpers', 'pytorch_ext',
               'vis', 'test', 'dataloaders')]:
        del sys.modules[m]


def run(cfg_name, H, W, n, expect_sd):
    cfg, bp, bc, _ = gg.load_reference(cfg_name)
    assert gg.sd_digest(bp.net.state_dict()) == expect_sd, 'weights differ from tests/golden/summary.json'
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for i in range(n):
            img = gg.make_image(i, H, W)
            p = os.path.join(tmp, 'i%d.l3c' % i)
            t0 = time.time()
            bpsp = bc.encode(img.long(), p)
            nbytes = os.path.getsize(p)
            os.remove(p)
            out.append(dict(img_seed=1000 + i, ref_bytes=nbytes, ref_bpsp=float(bpsp)))
            print(cfg_name, i, nbytes, '%.6f' % bpsp, '%.1fs' % (time.time() - t0), flush=True)
    return out


def main():
    summ = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'summary.json')))
    res = {'how': 'python oracle/gen_golden_batch.py (unmodified reference Bitcoding.encode, CPU, '
                  'torch %s, %d threads)' % (torch.__version__, torch.get_num_threads())}
    res['rgbs_256x256'] = run('cr_rgb_shared.cf', 256, 256, 32, summ['rgbs_sd_sha256'])
    _forget_reference()
    res['l3c_512x512'] = run('cr.cf', 512, 512, 16, summ['l3c_sd_sha256'])
    assert res['l3c_512x512'][0]['ref_bytes'] == summ['l3c_512x512_i0']['ref_bytes']
    assert res['rgbs_256x256'][0]['ref_bytes'] == summ['rgbs_256x256_i0']['ref_bytes']
    with open(os.path.join(ROOT, 'tests', 'golden', 'batch_bytes.json'), 'w') as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print('wrote tests/golden/batch_bytes.json')


if __name__ == '__main__':
    main()
