"""Bring-up helper: CUDA-event timing of the stages of one encode+decode batch."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import l3c_pytorch_b200 as l3c
from l3c_pytorch_b200 import config, engine as E

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    HW = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    prec = sys.argv[3] if len(sys.argv) > 3 else 'fp32'
    E.set_conv_precision(prec)
    torch.manual_seed(0)
    bp = l3c.MultiscaleBlueprint(config.ms_config('cr')).set_eval()
    codec = l3c.BatchCodec(bp)
    imgs = torch.stack([(torch.rand(3, HW, HW, generator=torch.Generator().manual_seed(1000 + i)) * 255).round().to(torch.uint8) for i in range(N)]).cuda()
    res = {}
    for it in range(3):
        torch.cuda.synchronize()
        t0 = ev(); out = bp.net(imgs); t1 = ev()
        datas, info = codec.encode_batch(imgs, out=out); t2 = ev()
        torch.cuda.synchronize(); w0 = time.perf_counter()
        S, _ = codec.decode_batch(datas, to_host=False); t3 = ev()
        torch.cuda.synchronize(); w1 = time.perf_counter()
        res = dict(forward_ms=t0.elapsed_time(t1), entropy_enc_ms=t1.elapsed_time(t2), decode_ms=t2.elapsed_time(t3), decode_wall_ms=(w1-w0)*1e3)
        ok = bool((S == imgs).all())
    px = N * HW * HW
    res.update(N=N, HW=HW, prec=prec, lossless=ok, bpsp=sum(map(len, datas)) * 8 / (3 * px),
               mpx_s=px / 1e6 / ((res['forward_ms'] + res['entropy_enc_ms'] + res['decode_ms']) / 1e3))
    print(json.dumps(res))

main()
