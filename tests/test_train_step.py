"""the PyTorch-autograd training step (l3c_pytorch_b200/train.py) against ONE step of the
unmodified reference (tests/golden/train_step.json, oracle/gen_golden_train.py): loss, per-scale bpsp, gradient
norms and the parameters after one RMSprop step.  Pure torch operators: runs on the CPU."""
import json
import os

import numpy as np
import torch

from tests import util


def test_training_step_matches_the_reference():
    from l3c_pytorch_b200 import MultiscaleBlueprint, config, train
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'train_step.json')) as f:
        g = json.load(f)
    torch.manual_seed(0)
    bp = MultiscaleBlueprint(config.ms_config('cr'), device='cpu')          # seed-0 default init = the reference's
    bp.net.float().cpu()
    bp.net.train()
    imgs = torch.stack([util.make_image(i, 32, 32) for i in range(2)]).float()
    opt = train.make_optimizer(bp)
    assert type(opt).__name__ == 'RMSprop' and opt.defaults['lr'] == 1e-4
    bp.net.zero_grad()
    loss_pc, bpsps = train.training_loss(bp, imgs)
    assert abs(float(loss_pc) - g['loss_pc']) < 2e-4 * g['loss_pc']
    np.testing.assert_allclose(bpsps, g['nonrecursive_bpsps'], rtol=2e-4)
    loss_pc.backward()
    params = dict(bp.net.named_parameters())
    for k, want in g['grad_norms'].items():
        got = float(params[k].grad.norm())
        assert abs(got - want) < 2e-3 * want + 1e-7, (k, got, want)
    total = float(torch.sqrt(sum((p.grad ** 2).sum() for p in bp.net.parameters() if p.grad is not None)))
    assert abs(total - g['total_grad_norm']) < 2e-3 * g['total_grad_norm']
    opt.step()
    for k, want in g['params_after_step'].items():
        np.testing.assert_allclose(params[k].detach().flatten()[:4].numpy(), want, rtol=1e-3, atol=2e-5)
    # one more step through the public entry point lowers nothing by itself, but must run and return floats
    l2, b2 = train.train_step(bp, opt, imgs)
    assert np.isfinite(l2) and len(b2) == 4
    # levels are not trained (requires_grad=False in the reference: net.py:123-127)
    assert bp.net.nets[0].enc.levels.grad is None
