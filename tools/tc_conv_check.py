"""Bring-up check of the tcgen05 conv against the fp32 FFMA conv (same inputs)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from l3c_pytorch_b200 import engine as E, _lib
from l3c_pytorch_b200.network import default_conv

def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps

torch.manual_seed(0)
for (N, H, W, cout, rate, kw) in [(1, 8, 16, 64, 1, {}), (2, 24, 40, 64, 1, {}), (2, 20, 33, 64, 2, {}), (1, 19, 37, 64, 4, {}),
                                   (2, 16, 16, 256, 1, dict(pixel_shuffle=True)), (2, 32, 32, 64, 1, dict(relu=True)),
                                   (2, 32, 48, 64, 1, dict(residual=True)), (16, 256, 256, 64, 1, {})]:
    conv = default_conv(64, cout, 3, rate=rate).cuda()
    x = torch.randn(N, H, W, 64, device='cuda')
    res = torch.randn(N, H, W, cout, device='cuda') if kw.get('residual') else None
    kw2 = {k: v for k, v in kw.items() if k != 'residual'}
    ref = E.conv2d(conv, x, residual=res, precision=_lib.PREC_FP32, **kw2)
    got = E.conv2d(conv, x, residual=res, precision=_lib.PREC_TF32, **kw2)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item(); scale = ref.abs().max().item()
    line = dict(N=N, H=H, W=W, cout=cout, rate=rate, kw=list(kw), max_abs_err=err, ref_max=scale, rel=err / scale)
    if N == 16:
        flops = 2 * 9 * 64 * cout * N * H * W
        tf = t(lambda: E.conv2d(conv, x, precision=_lib.PREC_TF32)); ff = t(lambda: E.conv2d(conv, x, precision=_lib.PREC_FP32))
        line.update(tc_ms=tf, tc_tflops=flops / tf / 1e9, ffma_ms=ff, ffma_tflops=flops / ff / 1e9)
    print(json.dumps(line), flush=True)
