"""oracle/gen_golden_train.py -- TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):
    python oracle/gen_golden_train.py

Golden for SURVEY.md section 8 row f4: ONE training step of the unmodified reference (train-mode forward with the
soft quantiser, DMLL loss, backward; multiscale_trainer.py:173-200) on the CPU with seed-0 default-init L3C
weights and a seeded 2 x 3x32x32 batch: loss_pc, per-scale bpsp, gradient norms of a few parameters, and the
parameter values after one RMSprop step -> tests/golden/train_step.json.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402

import torch  # noqa: E402

PROBE = ['heads.0.head.1.head.weight', 'nets.0.enc.down.weight', 'nets.1.enc.to_q.0.weight', 'nets.2.dec.tail.0.weight',
         'prob_clfs.0.atrous.lin.weight', 'prob_clfs.2.atrous.atrous.1.bias', 'nets.0.dec.body.3.body.0.weight']


def main():
    torch.set_grad_enabled(True)
    cfg, bp, bc, _ = gg.load_reference('cr.cf')
    bp.net.train()
    bp.losses.loss_dmol_rgb.train()
    bp.losses.loss_dmol_n.train()
    imgs = torch.stack([gg.make_image(i, 32, 32) for i in range(2)]).float()
    optim = torch.optim.RMSprop(bp.net.parameters(), cfg.lr.initial, weight_decay=cfg.weight_decay)
    bp.net.zero_grad()
    out = bp.forward(imgs)
    loss = bp.get_loss(out)
    loss.loss_pc.backward()
    params = dict(bp.net.named_parameters())
    grads = {k: float(params[k].grad.norm()) for k in PROBE}
    total = float(torch.sqrt(sum((p.grad ** 2).sum() for p in bp.net.parameters() if p.grad is not None)))
    optim.step()
    after = {k: [float(x) for x in params[k].detach().flatten()[:4]] for k in PROBE}
    res = {'config': 'cr.cf', 'batch': '2 x 3x32x32, image seeds 1000, 1001', 'loss_pc': float(loss.loss_pc),
           'nonrecursive_bpsps': [float(x) for x in loss.nonrecursive_bpsps], 'grad_norms': grads,
           'total_grad_norm': total, 'params_after_step': after, 'optim': 'RMSprop lr 1e-4'}
    print(res)
    with open(os.path.join(gg.ROOT, 'tests', 'golden', 'train_step.json'), 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
