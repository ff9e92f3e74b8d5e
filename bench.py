#!/usr/bin/env python
"""bench.py -- headline benchmark of the L3C encode/decode hot path (see BASELINE.json).

  python bench.py --gpus N --steps K --warmup W                 (our sm_100a path, one process per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W (reference CPU path, host cores)

One "step" = one lossless round trip (encode -> decode) of one batch of synthetic 3x512x512 images
per GPU (BASELINE.json configs[1]: 16 images per GPU, L3C cr.cf, seed-0 default-init weights).
Prints ONE JSON line (rank 0).  `value` = Mpixels/s with inputs resident in HBM; `e2e` = the same
round trip through the public `Bitcoding.encode_batch[_begin]/decode_batch` API from pinned host
buffers (H2D of the images and the containers, D2H of the containers and the decoded images inside
the timed region).  By default the K steps are software-pipelined (the encode of batch k+1 runs beside
the latency-bound decode of batch k); the strictly sequential figures are reported beside them
(`sequential`, `e2e.sequential_value`) and `--no-pipeline` makes them the headline.  Timing: CUDA events
on the stream all work forks from and joins, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'Mpixels/s encode+decode (lossless round-trip)'
IMAGES_PER_GPU = 16
HW = 512
CONV_FLOP_PER_PX_ROUNDTRIP = 2.230e6      # 1.368 (encode forward) + 0.863 (decode)


def make_images(first, n, hw=HW):
    import torch
    out = []
    for i in range(first, first + n):
        g = torch.Generator().manual_seed(1000 + i)
        out.append((torch.rand(3, hw, hw, generator=g) * 255).round().to(torch.uint8))
    return torch.stack(out)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for (t, line) in self.lines:
            if t < t0 or t > t1 + 0.3:
                continue
            f = [x.strip() for x in line.split(',')]
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except (ValueError, IndexError):
                continue
            for name, val in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'],
                                 f[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_roundtrip_mpx_s(n_images, hw, first_image=0):
    """Round trip of `n_images` hw x hw images with the reference's CPU algorithm as restated in
    oracle/model.py (byte-identical to the unmodified reference on the golden fixtures): PyTorch
    fp32 CPU convs (all host threads), PyTorch CPU CDF tables (torchac.py:174-213), the reference's
    own coder loop.  Returns (Mpx/s, seconds, bpsp, threads)."""
    import torch
    from l3c_pytorch_b200 import config
    from l3c_pytorch_b200.blueprint import MultiscaleBlueprint
    from oracle import model as om
    try:        # torchrun pins OMP_NUM_THREADS=1: give the CPU reference the physical cores it may use;
        # otherwise keep PyTorch's own default (measured faster than one thread per SMT sibling)
        if torch.get_num_threads() == 1:
            torch.set_num_threads(max(1, len(os.sched_getaffinity(0)) // 2))
    except (AttributeError, RuntimeError):
        pass
    torch.manual_seed(0)
    bp = MultiscaleBlueprint(config.ms_config('cr'), device='cpu')
    sd = {k: v.detach().cpu() for k, v in bp.net.state_dict().items()}
    imgs = make_images(first_image, n_images, hw)
    t0 = time.perf_counter()
    total_bytes = 0
    with torch.no_grad():
        for i in range(n_images):
            data = om.encode_image(sd, om.CFG_L3C, imgs[i], 'torch')
            dec = om.decode_image(sd, om.CFG_L3C, data, 'torch')
            assert bool((dec[0] == imgs[i].long()).all()), 'CPU reference round trip not lossless'
            total_bytes += len(data)
    dt = time.perf_counter() - t0
    px = n_images * hw * hw
    return px / 1e6 / dt, dt, total_bytes * 8 / (3 * px), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import torch
    steps, warm = args.steps, args.warmup
    hw = HW if (steps + warm) <= 6 else 256          # keep the whole run within a few minutes
    for _ in range(warm):
        cpu_roundtrip_mpx_s(1, 128)
    vals = []
    t_all = 0.0
    for s in range(steps):
        v, dt, bpsp, thr = cpu_roundtrip_mpx_s(1, hw, first_image=s)
        vals.append(v)
        t_all += dt
    px = steps * hw * hw
    value = px / 1e6 / t_all
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'Mpixels/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': 1e3 * t_all / steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'L3C cr.cf, %d x 3x512x512 per GPU, encode+decode round trip' % IMAGES_PER_GPU,
                   'sample': '1 image of 3x%dx%d per step on the host CPU' % (hw, hw)},
        'cpu_baseline': {'value': value, 'unit': 'Mpixels/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': '%d round trip(s) of one 3x%dx%d image (oracle/model.py, PyTorch-CPU CDF path, '
                                   'byte-identical to the unmodified reference on tests/golden)' % (steps, hw, hw)},
        'e2e': {'value': value, 'unit': 'Mpixels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'bpsp': bpsp,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import l3c_pytorch_b200 as l3c
    from l3c_pytorch_b200 import config, dist as l3c_dist, engine as E, _lib

    rank, world, local_rank = l3c_dist.init_from_env()
    assert torch.cuda.is_available(), 'bench.py (our arm) needs a GPU; there is no CPU fallback'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    E.set_conv_precision(args.precision)
    torch.manual_seed(0)
    bp = l3c.MultiscaleBlueprint(config.ms_config('cr'), device=dev).set_eval()
    bc = l3c.Bitcoding(bp)
    codec = bc.codec
    n_img = args.images_per_gpu
    n_global = n_img * world
    lo, hi = l3c_dist.shard_bounds(n_global, rank, world)
    # a few distinct batches so consecutive steps do not see identical inputs; activations are
    # ~1 GB per layer at this batch size, far beyond the 126 MB L2
    n_sets = 2
    host_sets = [make_images(lo + s * n_global, n_img).pin_memory() for s in range(n_sets)]
    dev_sets = [h.to(dev) for h in host_sets]
    launches = {'n': 0}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def shapes_of(info):
        return [(C, H, W) for (_, C, H, W) in info['shapes']]

    def step_resident(imgs):
        blob, info = codec.encode_batch(imgs, to_host=False)
        S = codec.decode_device(blob, info['stream_offsets'], info['lens'], shapes_of(info))
        return S, info

    def step_e2e(host_imgs):
        datas, bpsps = bc.encode_batch(host_imgs)
        dec = bc.decode_batch(datas)
        back = torch.cat(dec, 0).to(torch.uint8).cpu()        # D2H of the result
        return back, datas

    # Software pipelining over steps (the default): the decode of a batch is bound by the serial range
    # decoder and leaves most SMs idle, so the encode of the NEXT batch runs beside it, on a stream that
    # is confined to the SMs the decoders do not own.  K complete round trips -- including the
    # un-overlapped first encode and last decode -- lie inside the timed region.
    side_stream = codec.encode_stream(dev, 3 * n_img)
    # the decode is the latency-critical half: it runs on a high-priority stream so that its kernels
    # are scheduled ahead of the encode's whenever both are waiting for SMs
    main_stream = torch.cuda.Stream(device=dev, priority=-1) if args.pipeline else torch.cuda.current_stream()

    def run_resident(steps, first_set=0):
        if not args.pipeline:
            for s in range(steps):
                S, info = step_resident(dev_sets[(first_set + s) % n_sets])
            return S, info
        dbg = os.environ.get('L3C_BENCH_DEBUG')
        cur = torch.cuda.current_stream()
        main_stream.wait_stream(cur)              # the timing events live on `cur`: fork from it ...
        side_stream.wait_stream(cur)
        with torch.cuda.stream(side_stream):
            job = codec.encode_begin(dev_sets[first_set % n_sets])
        for s in range(steps):
            t0 = time.perf_counter()
            blob, info = job.finish(to_host=False)
            t1 = time.perf_counter()
            job = None
            if s + 1 < steps:
                with torch.cuda.stream(side_stream):
                    job = codec.encode_begin(dev_sets[(first_set + s + 1) % n_sets])
            t2 = time.perf_counter()
            with torch.cuda.stream(main_stream):
                main_stream.wait_event(info['ready'])
                blob.record_stream(main_stream)
                S = codec.decode_device(blob, info['stream_offsets'], info['lens'], shapes_of(info))
            if dbg:
                print('pipelined step %d: finish %.1f ms, begin(next) %.1f ms, decode issue %.1f ms'
                      % (s, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (time.perf_counter() - t2)), file=sys.stderr)
        cur.wait_stream(main_stream)              # ... and join it again
        cur.wait_stream(side_stream)
        return S, info

    def run_e2e(steps, first_set=0):
        if not args.pipeline:
            for s in range(steps):
                back, datas = step_e2e(host_sets[(first_set + s) % n_sets])
            return back, datas
        cur = torch.cuda.current_stream()
        main_stream.wait_stream(cur)
        side_stream.wait_stream(cur)
        job = bc.encode_batch_begin(host_sets[first_set % n_sets], stream=side_stream)
        # the decoded images come back through two pinned buffers; the read-back of step s is awaited
        # after step s+1 has been issued, so the host prepares the next decode while this one runs
        bufs = [torch.empty_like(host_sets[0]).pin_memory() for _ in range(2)]
        pending = None
        back = None
        for s in range(steps):
            datas, _ = job.finish()
            job = None
            if s + 1 < steps:
                job = bc.encode_batch_begin(host_sets[(first_set + s + 1) % n_sets], stream=side_stream)
            with torch.cuda.stream(main_stream):
                dec = bc.decode_batch(datas)                   # containers -> GPU -> images (asynchronous)
                bufs[s % 2].copy_(torch.cat(dec, 0).to(torch.uint8), non_blocking=True)    # D2H of the result
                ev = torch.cuda.Event()
                ev.record(main_stream)
            if pending is not None:
                pending.synchronize()
            pending = ev
        pending.synchronize()
        back = bufs[(steps - 1) % 2]
        cur.wait_stream(main_stream)
        cur.wait_stream(side_stream)
        return back, datas

    # ---- warm-up + correctness (outside the timed region)
    for w in range(max(args.warmup, 1)):
        S, info = step_resident(dev_sets[w % n_sets])
    assert torch.equal(S, dev_sets[(max(args.warmup, 1) - 1) % n_sets]), 'round trip is not lossless'
    if args.pipeline:                      # the side stream has its own allocator pool: warm it up too
        S, info = run_resident(max(args.warmup, 2))
        assert torch.equal(S, dev_sets[(max(args.warmup, 2) - 1) % n_sets]), 'pipelined round trip is not lossless'
    sizes = info['sizes']
    counts = l3c_dist.gather_byte_counts(sizes, n_global, rank, world)       # the one collective (NCCL)
    bpsp = l3c_dist.global_bpsp(counts, 3 * HW * HW)
    # parity with the reference's own torchac path on the same weights/image (golden: image seed 1000)
    parity = None
    try:
        with open(os.path.join(ROOT, 'tests', 'golden', 'summary.json')) as f:
            ref0 = json.load(f)['l3c_512x512_i0']
        d0, _ = codec.encode_batch(dev_sets[0][:1])
        if rank == 0:
            parity = {'image': 'seed 1000, 3x512x512', 'bytes': len(d0[0]), 'reference_bytes': ref0['ref_bytes'],
                      'abs_dbpsp': abs(len(d0[0]) - ref0['ref_bytes']) * 8 / (3.0 * HW * HW)}
    except (OSError, KeyError):
        pass

    # ---- timed: device-resident
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launches0 = E.LAUNCHES['n']
    S, info = run_resident(args.steps)
    launches_timed = E.LAUNCHES['n'] - launches0
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = float(ms)
    px_step_global = n_global * HW * HW
    value = px_step_global * args.steps / 1e6 / (ms_total / 1e3)
    assert torch.equal(S, dev_sets[(args.steps - 1) % n_sets]), 'timed round trip is not lossless'

    def timed_sequential(fn):
        """the same K steps strictly one after the other (reported beside the pipelined numbers)"""
        barrier()
        tw = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(args.steps):
            fn(s % n_sets)
        b.record()
        barrier()
        t = torch.tensor([max(a.elapsed_time(b), 0.0), (time.perf_counter() - tw) * 1e3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    seq_value = seq_e2e = None
    if args.pipeline:
        seq_ms, _ = timed_sequential(lambda i: step_resident(dev_sets[i]))
        seq_value = px_step_global * args.steps / 1e6 / (seq_ms / 1e3)

    # ---- timed: end to end through the public API (host buffers, copies inside)
    back, datas = step_e2e(host_sets[0])
    assert torch.equal(back, host_sets[0]), 'e2e round trip is not lossless'
    if args.pipeline:
        run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    back, datas = run_e2e(args.steps)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms2 = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = px_step_global * args.steps / 1e6 / (float(ms2) / 1e3)
    assert torch.equal(back, host_sets[(args.steps - 1) % n_sets]), 'timed e2e round trip is not lossless'
    if args.pipeline:
        seq_ev, seq_wall = timed_sequential(lambda i: step_e2e(host_sets[i]))
        seq_e2e = px_step_global * args.steps / 1e6 / (max(seq_ev, seq_wall) / 1e3)
    cont_bytes = sum(len(d) for d in datas)
    img_bytes = n_img * 3 * HW * HW

    # ---- roofline of the dominant kernel: the 3x3 64->64 convolution at 256x256 (34 of the ~120
    #      conv launches of a round trip and ~60 % of its FLOPs run on exactly this shape)
    conv = bp.net.nets[0].enc.body[0].body[0]
    x = torch.randn(n_img, HW // 2, HW // 2, 64, device=dev)
    y = torch.empty_like(x)
    for _ in range(3):
        E.conv2d(conv, x, out=y)
    torch.cuda.synchronize()
    reps = 10
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(reps):
        E.conv2d(conv, x, out=y)
    c1.record()
    torch.cuda.synchronize()
    conv_ms = c0.elapsed_time(c1) / reps
    conv_flops = 2.0 * 9 * 64 * 64 * n_img * (HW // 2) ** 2
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak_tf = peaks.get('bf16_tflops', 1590.0)
    achieved_tf = conv_flops / (conv_ms / 1e3) / 1e12
    roofline = {'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': achieved_tf / peak_tf,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this shape, one launch, from the
                # ncu --set full capture committed as profiles/r01b_conv3x3_tcgen05_v2_ncu_full_raw.csv
                # (algorithmic: 268 MB read + 268 MB written)
                'traffic': 484.4e6 if (args.precision == 'tf32' and n_img == 16) else None,
                'kernel': 'conv3x3 64->64, %dx256x256 NHWC fp32 (%s path)' % (n_img, args.precision),
                'peak_source': 'MEASURED_PEAKS.json bf16 burst' if peaks else 'fallback 1.59 PFLOP/s',
                'whole_step_conv_tflops': CONV_FLOP_PER_PX_ROUNDTRIP * n_img * HW * HW * args.steps /
                (ms_total / 1e3) / 1e12}

    # ---- where the step goes (one extra, untimed-for-the-metric round trip with CUDA events): the
    #      dominant kernel by time is the serial range decoder, which is latency-bound (one warp per
    #      stream, ~180 ns per symbol whatever the number of streams), not HBM- or tensor-bound
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    ev[0].record()
    out_net = bp.net(dev_sets[0])
    ev[1].record()
    blob, info = codec.encode_batch(dev_sets[0], out=out_net, to_host=False)
    ev[2].record()
    codec.decode_device(blob, info['stream_offsets'], info['lens'], [(C, H, W) for (_, C, H, W) in info['shapes']])
    ev[3].record()
    torch.cuda.synchronize()
    n_sym_rgb = HW * HW
    dec_ms = ev[2].elapsed_time(ev[3])
    breakdown = {'forward_ms': ev[0].elapsed_time(ev[1]), 'entropy_encode_ms': ev[1].elapsed_time(ev[2]),
                 'decode_ms': dec_ms,
                 'serial_symbols_per_stream': {'rgb': n_sym_rgb, 'z1': n_sym_rgb // 4, 'z2': n_sym_rgb // 16,
                                               'z3': n_sym_rgb // 64},
                 'streams_in_flight': n_img * 18,
                 'decoder_algorithmic_GBps': n_img * 3 * n_sym_rgb * 515e-9 / (dec_ms / 1e3),
                 'note': 'range coder = one warp per stream, bounded by dependent-issue latency per symbol; '
                         'HBM use of the decoder (512 B CDF row + code + symbol per symbol) is ~1 % of peak'}

    # ---- CPU baseline beside it (rank 0, N=1 only, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cb, thr = cpu_roundtrip_mpx_s(1, HW)
        cpu = {'value': v, 'unit': 'Mpixels/s', 'cores': thr, 'kind': 'port',
               'sample': '1 round trip of one 3x512x512 image (%.1f s): oracle/model.py with the reference\'s '
                         'PyTorch-CPU CDF path (byte-identical to the unmodified reference on tests/golden)' % dt,
               'bpsp': cb}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'Mpixels/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'L3C cr.cf (3 scales, seed-0 default init), %d x 3x512x512 uint8 noise images '
                                   'per GPU, encode+decode round trip, byte-compatible .l3c containers' % n_img,
                       'global_batch': n_global, 'parallelism': 'images sharded over %d GPU(s), no data-path '
                                                                'collective' % world,
                       'conv_precision': args.precision,
                       'pipelining': ('encode of batch k+1 overlaps decode of batch k (separate SM partitions); all '
                                      '%d round trips, incl. the un-overlapped first encode and last decode, are '
                                      'inside the timed region; "sequential" = the same steps one after the other'
                                      % args.steps) if args.pipeline else 'none (sequential steps)',
                       'l2': 'working set >> L2 (1 GB of activations per layer), inputs alternate between batches'},
            'bpsp': bpsp, 'bpsp_parity': parity,
            'e2e': {'value': e2e_value, 'unit': 'Mpixels/s', 'h2d_bytes_per_step': img_bytes + cont_bytes,
                    'd2h_bytes_per_step': cont_bytes + img_bytes, 'sequential_value': seq_e2e},
            'sequential': {'value': seq_value, 'unit': 'Mpixels/s'},
            'gpu_launches': None,
            'clocks': clocks,
            'roofline': roofline,
            'breakdown': breakdown,
            'cpu_baseline': cpu,
        }
        # kernels of libl3c_b200.so launched inside the timed (device-resident) region, counted at the
        # C-ABI call sites (engine.LAUNCHES)
        line['gpu_launches'] = launches_timed
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('L3C_CONV_PRECISION', 'tf32'),
                    choices=['fp32', 'tf32', 'tf32x3', 'bf16'])
    ap.add_argument('--images-per-gpu', type=int, default=IMAGES_PER_GPU)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-pipeline', dest='pipeline', action='store_false',
                    help='strictly sequential steps: encode(k), decode(k), encode(k+1), ...')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
