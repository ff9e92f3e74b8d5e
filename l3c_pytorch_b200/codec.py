"""Batched, device-resident encode/decode of `.l3c` containers -- the B200-first core that the
reference-shaped `Bitcoding` class (bitcoding.py) drives.

What the reference does one image, one scale, one channel at a time with a GPU->CPU round trip per
channel (/root/reference/src/bitcoding/bitcoding.py:85-123,163-294, coders.py:38-90) is done here
for a whole batch of equally sized images with
  * ONE network pass,
  * one interval kernel per scale (no CDF table on the encode side),
  * ONE range-coder launch covering every stream of every image (N * 18 warps for L3C),
  * ONE gather kernel + ONE device->host copy producing the byte-exact container layout
    (bitcoding.py:326-363: 4 x u16 padding | per scale: u8 C, u16 H, u16 W | per channel: u32 len +
    stream | 4-byte magic).
Decoding uploads the concatenated containers once and decodes every stream in place.
"""
import os
import struct

import numpy as np
import torch

from . import _lib
from . import engine as E
from .dmll import non_shared_get_K

MAGIC = b'\x46\xE2\x84\x92'          # bitcoding.py:36


def _on_device(fn):
    """Run a codec entry point with the blueprint's GPU as the current CUDA device: the kernels of the
    shared library launch on whatever device is current, and `current_stream()` is per device -- a
    Blueprint(device='cuda:1') must not depend on the caller having called torch.cuda.set_device(1)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        dev = self.codec.blueprint.device if isinstance(self, EncodeJob) else self.blueprint.device
        if dev.type != 'cuda':
            return fn(self, *a, **kw)
        with torch.cuda.device(dev):
            return fn(self, *a, **kw)
    return wrapped


def _under_injection_tool():
    """True when a CUDA injection tool (Nsight Compute / Systems) is attached to this process: either it
    announced itself through the environment or its injection libraries are mapped.  Nsight Compute cannot
    profile kernels launched into green-context streams (it loses the process at the first one:
    "Failed to prepare kernel for profiling"), so the SM partitions are switched off under it -- same
    kernels, same bytes, ordinary streams; L3C_SM_PARTITION=1 forces them on."""
    if any(k in os.environ for k in ('CUDA_INJECTION64_PATH', 'CUDA_INJECTION32_PATH', 'NVTX_INJECTION64_PATH')):
        return True
    try:
        with open('/proc/self/maps') as f:
            for line in f:
                low = line.lower()
                if 'nsight' in low or 'injection' in low:
                    return True
    except OSError:
        pass
    return False


def _slot_cap(n_sym):
    """worst case: every symbol has probability 2^-16 -> 17 bits, plus termination."""
    return ((n_sym * 17 + 7) // 8 + 64 + 3) & ~3


TILED_MAGIC = b'L3CT'                # tiled containers: this tag + u16 tile_h + u16 tile_w follow the padding tuple


def tile_grid(H, W, tile):
    """[(offset, n_sym)] of the streams of one H x W symbol plane in stream order: the whole plane (compat
    format) or its tiles of tile = (th, tw) in row-major tile order (tile-major symbol order, dmll.cu)."""
    if tile is None:
        return [(0, H * W)]
    th, tw = tile
    out, off = [], 0
    for ty in range(-(-H // th)):
        h = min(th, H - ty * th)
        for tx in range(-(-W // tw)):
            w = min(tw, W - tx * tw)
            out.append((off, h * w))
            off += h * w
    return out


class ContainerLayout(object):
    """Byte layout of one container given its per-stream lengths (scale order coarse -> fine).
    tile = None: the reference's `.l3c` layout (bitcoding.py:326-363), one stream per channel and scale.
    tile = (th, tw): the throughput layout -- every channel plane is cut into tiles that are coded as
    independent streams (thousands per batch instead of 18 per image); NOT readable by the reference and a few
    bytes per tile larger (4-byte length + termination), reported by bench.py as `tiled`."""

    def __init__(self, shapes, tile=None):
        self.shapes = shapes            # [(scale, C, H, W)] coarse -> fine
        self.tile = tile

    def header_and_offsets(self, lens, pad_tuple):
        """lens: stream byte counts in container order.  Returns (total size, [(offset, bytes)]
        header pieces, [stream offsets])."""
        pieces = [(0, struct.pack('<4H', *pad_tuple))]
        pos = 8
        if self.tile is not None:
            pieces.append((pos, TILED_MAGIC + struct.pack('<HH', *self.tile)))
            pos += 8
        offs = []
        i = 0
        for (_, C, H, W) in self.shapes:
            pieces.append((pos, struct.pack('<BHH', C, H, W)))
            pos += 5
            for _ in range(C * len(tile_grid(H, W, self.tile))):
                n = int(lens[i])
                pieces.append((pos, struct.pack('<I', n)))
                pos += 4
                offs.append(pos)
                pos += n
                i += 1
            pieces.append((pos, MAGIC))
            pos += 4
        return pos, pieces, offs


def container_tile(data):
    """tile size (th, tw) of a tiled container, None for the reference layout."""
    if len(data) >= 16 and bytes(data[8:12]) == TILED_MAGIC:
        return struct.unpack_from('<HH', data, 12)
    return None


def parse_container(data):
    """-> (pad_tuple, [(C, H, W, [(offset, length), ...])] coarse -> fine); tiled containers (container_tile)
    list C * tiles streams per scale.  Raises ValueError on a malformed file (the reference asserts on the
    magic, bitcoding.py:154)."""
    if len(data) < 8:
        raise ValueError('container too short')
    pad_tuple = struct.unpack_from('<4H', data, 0)
    pos = 8
    tile = container_tile(data)
    if tile is not None:
        if tile[0] < 1 or tile[1] < 1:
            raise ValueError('tiled container with an empty tile size')
        pos = 16
    scales = []
    while pos < len(data):
        if pos + 5 > len(data):
            raise ValueError('truncated scale header')
        C, H, W = struct.unpack_from('<BHH', data, pos)
        pos += 5
        streams = []
        for _ in range(C * len(tile_grid(H, W, tile))):
            if pos + 4 > len(data):
                raise ValueError('truncated stream length')
            n, = struct.unpack_from('<I', data, pos)
            pos += 4
            if pos + n > len(data):
                raise ValueError('stream runs past the end of the file')
            streams.append((pos, n))
            pos += n
        if data[pos:pos + 4] != MAGIC:
            raise ValueError('scale separator missing: not a valid .l3c file')
        pos += 4
        scales.append((C, H, W, streams))
    return pad_tuple, scales


class DecodeLane(object):
    """Streams of one decode in flight (BatchCodec.lanes)."""
    __slots__ = ('main', 'dec', 'bld', 'partitioned')


class EncodeJob:
    """An encode in flight (BatchCodec.encode_begin).  finish() waits for the stream lengths, lays the
    containers out, gathers the streams into one blob on the job's stream and (to_host) brings the
    bytes back: same results as encode_batch."""

    @_on_device
    def finish(self, to_host=True):
        codec, N, per_img, caps, shapes = self.codec, self.N, self.per_img, self.caps, self.shapes
        self.coded.synchronize()                                                   # sync #1 (tiny)
        lens = self.lens_host.numpy().astype(np.int64).reshape(N, per_img)
        if (lens > caps[None, :]).any():
            raise RuntimeError('range coder output exceeded its slot (corrupt CDF?)')
        with torch.cuda.stream(self.stream):
            # ---- container layout + gather + single D2H
            layout = ContainerLayout(shapes, self.tile)
            sizes, pieces_all, dst = [], [], np.zeros((N, per_img), np.int64)
            pos = 0
            starts = []
            for n in range(N):
                total, pieces, offs = layout.header_and_offsets(lens[n], self.pad_tuple)
                starts.append(pos)
                dst[n] = pos + np.asarray(offs, np.int64)
                pieces_all.append(pieces)
                sizes.append(total)
                pos += (total + 15) & ~15
            blob = torch.empty(pos + 16, dtype=torch.uint8, device=self.dev)
            E.pack_streams(self.desc_dev, self.lens_dev, dst.reshape(-1), N * per_img, blob)
            info = dict(sizes=sizes, starts=starts, lens=lens, shapes=shapes, out=self.out, stream_offsets=dst,
                        tile=self.tile)
            if not to_host:
                self.packed = torch.cuda.Event()
                self.packed.record(self.stream)
                info['ready'] = self.packed          # consumers on another stream wait for this event
                self.keep = None
                return blob, info
            host = torch.empty(blob.shape, dtype=torch.uint8, pin_memory=True)
            host.copy_(blob, non_blocking=True)
            self.stream.synchronize()                                              # sync #2
        self.keep = None
        buf = host.numpy()
        datas = []
        for n in range(N):
            s0 = starts[n]
            for (o, b) in pieces_all[n]:
                buf[s0 + o:s0 + o + len(b)] = np.frombuffer(b, np.uint8)
            datas.append(buf[s0:s0 + sizes[n]].tobytes())
        return datas, info


class BatchCodec(object):
    def __init__(self, blueprint):
        self.blueprint = blueprint
        self.net = blueprint.net
        self._const = {}

    # ------------------------------------------------------------------------------------------
    def iter_scale_dmll(self):
        """coarse -> fine: (scale, dmll, uniform) -- bitcoding.py:163-169."""
        losses = self.blueprint.losses
        for scale in reversed(range(self.net.scales + 1)):
            yield (scale, losses.loss_dmol_rgb if scale == 0 else losses.loss_dmol_n,
                   scale == self.net.scales)

    def _uniform(self, L, device):
        key = ('uniform', L, str(device))
        if key not in self._const:
            row = E.uniform_cdf_row(L).astype(np.int64)
            hi = np.concatenate([row[1:L], [65536]])
            lut = np.zeros(256, np.int64)
            lut[:L] = row[:L] | ((hi - 1) << 16)
            lut_dev = torch.from_numpy(lut.astype(np.uint32).view(np.int32).copy()).to(device)
            pitch = E.table_pitch(L)
            trow = np.zeros(pitch, np.uint16)
            trow[:L] = row[:L]
            row_dev = torch.from_numpy(trow.view(np.int16).copy()).to(device)
            self._const[key] = (lut_dev, row_dev)
        return self._const[key]

    def _rgb_shift(self, device):
        """RGB baselines feed `S - 255*rgb_mean` to the decoder nets (net.py:77-80; SURVEY finding 2)."""
        if not self.net._rgb:
            return None
        return self.net.nets[0].enc._consts(device)[1]

    # ------------------------------------------------------------------------------------------
    def encode_batch(self, imgs_u8, pad_tuple=(0, 0, 0, 0), out=None, to_host=True, tile=None):
        """imgs_u8: uint8 [N,3,H,W] on the GPU, H and W multiples of 2**num_scales.
        Returns a list of N container byte strings (or, with to_host=False, the device blob,
        per-image sizes and offsets, leaving the bytes in HBM).  `out`: a precomputed network Out.
        tile = (th, tw): tiled containers (ContainerLayout)."""
        return self.encode_begin(imgs_u8, pad_tuple, out, tile).finish(to_host)

    @_on_device
    def encode_begin(self, imgs_u8, pad_tuple=(0, 0, 0, 0), out=None, tile=None):
        """First half of encode_batch: enqueues the whole GPU side of an encode (networks, intervals,
        the range-coder launch, the copy of the stream lengths to pinned memory) on the CURRENT stream
        and returns without waiting.  EncodeJob.finish() completes it.  Lets a caller overlap the
        encode of one batch with other work, e.g. the latency-bound decode of the previous batch."""
        assert imgs_u8.dtype == torch.uint8 and imgs_u8.dim() == 4 and imgs_u8.shape[1] == 3
        dev = imgs_u8.device
        N = imgs_u8.shape[0]
        if out is None:
            out = self.net(imgs_u8, intervals_of=self.blueprint.losses)      # f16 mode: DMLL head fused into the convs
        K = self.net.config_ms.prob.K

        # ---- per-symbol intervals, scale by scale (coarse -> fine = container order)
        ivs, shapes = [], []
        for scale, dmll, uniform in self.iter_scale_dmll():
            S = out.S_u8[scale]
            _, C, H, W = S.shape
            shapes.append((scale, C, H, W))
            if uniform:
                ivs.append(E.lut_intervals(S, self._uniform(dmll.L, dev)[0]))
            elif out.IV is not None and out.IV[scale] is not None:
                ivs.append(out.IV[scale])                                    # produced by the fused head
            else:
                l = out.P_nhwc[scale]
                assert non_shared_get_K(l.shape[-1], C) == K
                ivs.append(E.dmll_intervals(l, S, dmll.targets(dev), C, K, dmll.L, dmll.rgb_scale))

        # ---- one range-coder launch for all streams; descriptor order = (image, scale, channel[, tile])
        if tile is not None:
            tile = (int(tile[0]), int(tile[1]))
            # intervals are produced in raster order; the streams of a tiled container read them tile by tile
            ivs = [E.reorder_tiles(iv, H, W, tile, True) for iv, (_, C, H, W) in zip(ivs, shapes)]
        streams = [(si, c, off, n) for si, (_, C, H, W) in enumerate(shapes) for c in range(C)
                   for (off, n) in tile_grid(H, W, tile)]
        per_img = len(streams)
        caps = np.array([_slot_cap(n) for (_, _, _, n) in streams], np.int64)
        slot_off = np.concatenate([[0], np.cumsum(caps)])
        img_slot_bytes = int(slot_off[-1])
        slots = torch.empty(N * img_slot_bytes, dtype=torch.uint8, device=dev)
        desc = np.zeros(N * per_img, dtype=_lib.ENC_STREAM_DTYPE)
        d = desc.reshape(N, per_img)
        bases = np.array([iv.data_ptr() for iv in ivs], np.int64)
        st = np.array(streams, np.int64)                                    # [per_img][4]
        CHW = np.array([(C, H * W) for (_, C, H, W) in shapes], np.int64)[st[:, 0]]      # per stream: C, HW of its scale
        img = np.arange(N, dtype=np.int64)[:, None]
        d['intervals'][:] = bases[st[:, 0]][None, :] + ((img * CHW[None, :, 0] + st[None, :, 1]) * CHW[None, :, 1]
                                                        + st[None, :, 2]) * 4
        d['n_sym'][:] = st[None, :, 3]
        d['out_cap'][:] = caps[None, :]
        d['out'][:] = slots.data_ptr() + img * img_slot_bytes + slot_off[None, :-1]
        desc_dev, lens_dev = E.ac_encode_streams(desc, dev)
        lens_host = torch.empty(lens_dev.shape, dtype=lens_dev.dtype, pin_memory=True)
        lens_host.copy_(lens_dev, non_blocking=True)
        job = EncodeJob()
        job.codec, job.stream = self, torch.cuda.current_stream()
        job.coded = torch.cuda.Event()
        job.coded.record(job.stream)
        job.keep = (imgs_u8, ivs, slots)                       # alive until finish()
        job.N, job.per_img, job.caps, job.shapes, job.out, job.tile = N, per_img, caps, shapes, out, tile
        job.pad_tuple, job.desc_dev, job.lens_dev, job.lens_host, job.dev = pad_tuple, desc_dev, lens_dev, lens_host, dev
        return job

    # ------------------------------------------------------------------------------------------
    @_on_device
    def decode_batch(self, datas, to_host=True, lane=None):
        """datas: list of container byte strings of equally shaped images.
        Returns (uint8 [N,3,H,W] incl. padding, pad_tuple list)."""
        dev = self.blueprint.device
        N = len(datas)
        parsed = [parse_container(d) for d in datas]
        ref_shapes = [(C, H, W) for (C, H, W, _) in parsed[0][1]]
        tile = container_tile(datas[0])
        for p, dta in zip(parsed, datas):
            if [(C, H, W) for (C, H, W, _) in p[1]] != ref_shapes or container_tile(dta) != tile:
                raise ValueError('decode_batch needs containers of identical shape')
        # one upload of everything
        starts = np.zeros(N, np.int64)
        pos = 0
        for n, dta in enumerate(datas):
            starts[n] = pos
            pos += (len(dta) + 15) & ~15
        host = torch.zeros(pos + 16, dtype=torch.uint8, pin_memory=True)
        hb = host.numpy()
        for n, dta in enumerate(datas):
            hb[starts[n]:starts[n] + len(dta)] = np.frombuffer(dta, np.uint8)
        blob = host.to(dev, non_blocking=True)
        offs = np.array([[starts[n] + o for (_, _, _, st) in parsed[n][1] for (o, _) in st] for n in range(N)],
                        np.int64)
        lens = np.array([[ln for (_, _, _, st) in parsed[n][1] for (_, ln) in st] for n in range(N)], np.int64)
        S = self.decode_device(blob, offs, lens, ref_shapes, lane=lane, tile=tile)
        pads = [p[0] for p in parsed]
        if to_host:
            return S.cpu(), pads
        return S, pads

    def _decode_rgb_pipelined(self, l, S, tg, C, K, L, table, d, dev, n_chunks=None, lane=None):
        """RGB scale: channel c's means depend on the decoded channels < c at the same pixel
        (logistic_mixture.py:262-272), so the reference codes R, G, B strictly one after the other.
        Here the three serial decoders run concurrently, staggered by one chunk of pixels: as soon as
        chunk j of channel c-1 is decoded, the CDF rows of chunk j of channel c are built (on the SM
        group the decoders do not own, _rgb_streams) and channel c's warps continue -- coder state is
        carried across launches.  Serial depth drops from 3*HW to (1 + 2/n_chunks)*HW symbols (whole
        decode at 16x512^2 with 32 / 64 / 128 chunks: 120.5 / 118.7 / 117.7 ms when measured)."""
        N, HW = S.shape[0], S.shape[2] * S.shape[3]
        if n_chunks is None:
            n_chunks = int(os.environ.get('L3C_RGB_CHUNKS', 64))
        csz = max(2048, -(-HW // n_chunks))
        csz = -(-csz // 64) * 64
        state = torch.zeros(N * C * 4, dtype=torch.int32, device=dev)
        d['state'][:] = state.data_ptr() + 16 * (np.arange(N)[:, None] * C + np.arange(C)[None, :])
        descs = [E._desc_to_device(np.ascontiguousarray(d[:, c]), dev) for c in range(C)]
        cur = torch.cuda.current_stream()
        if lane is None:
            lane = self.lanes(dev, N * C, 1)[0][0]
        # the chunk loop itself (2 x 3 x n_chunks launches + as many event records / waits) runs inside the library
        E.decode_rgb_pipelined(l, S, tg, K, L, table, descs, csz, cur, lane.bld, lane.dec)

    def encode_stream(self, dev, n_decoders, n_lanes=1):
        """A stream for encodes that run beside decodes (EncodeJob pipelining): confined to the SM group of
        the row builders, so that its persistent conv kernels never sit on the decoders' SMs."""
        return self.lanes(dev, n_decoders, n_lanes)[1]

    def lanes(self, dev, n_decoders, n_lanes=1):
        """-> ([DecodeLane] * n_lanes, encode stream).  A DecodeLane is the set of streams ONE decode in flight
        uses: `main` (decoder networks, CDF rows of the bottleneck scales), `bld[3]` (CDF-row builders of the
        R, G, B channels), `dec[3]` (the range decoders).  The `dec` streams of all lanes own a group of SMs
        (driver green contexts, l3c_partition_streams) sized for `L3C_DEC_WARPS_PER_SM` decoder-kernel warps
        per SM (default 4 for one lane = one latency-bound warp per SM sub-partition, 8 for several lanes; a
        decoder warp that shares its scheduler with throughput-bound warps runs ~1.5x slower); everything else -- including the encode
        stream, which has the lowest priority -- runs on the remaining SMs.  Several lanes = several batches
        being decoded side by side: the decode of ONE batch is bound by the serial chain of its range coder
        and keeps only ~1/6 of the GPU busy.  L3C_SM_PARTITION=0, a driver without green contexts or an
        attached profiler give ordinary streams (slower, same bytes)."""
        key = (str(dev), n_decoders, n_lanes)
        cache = self.__dict__.setdefault('_lane_cache', {})
        if key not in cache:
            part = None
            # encode streams: an encode is a throughput phase (the networks) followed by a latency-bound one (the
            # range encoders, ~19 ms on one warp per stream); encodes on the same stream run one after the other
            n_enc = max(1, int(os.environ.get('L3C_ENC_STREAMS', 2)))
            want_part = os.environ.get('L3C_SM_PARTITION', '1') != '0'
            if want_part and _under_injection_tool() and 'L3C_SM_PARTITION' not in os.environ:
                # Nsight Compute cannot profile kernels launched into green-context streams (it lost the
                # process at the first one): under a CUDA injection tool use ordinary streams
                import sys
                print('l3c_pytorch_b200: profiler injection detected, SM partitions off '
                      '(set L3C_SM_PARTITION=1 to force them)', file=sys.stderr)
                want_part = False
            if want_part:
                n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
                # one lane: 4 warps per SM (lowest latency); several lanes: 8, from four lanes on 12 (measured at
                # 16 x 512^2: 3 lanes on 36 / 48 / 72 SMs -> 86 / 69 / 63 Mpx/s, 4 lanes on 32 / 48 SMs -> 89 / 80: what
                # the decoders' group takes is lost to the throughput work)
                wps = int(os.environ.get('L3C_DEC_WARPS_PER_SM', 4 if n_lanes == 1 else (8 if n_lanes <= 3 else 12)))
                # every stream = decoder warp + helper warp (measured at 16 x 512^2, one lane, 4 warps per SM:
                # 48 / 24 / 16 SMs -> decode 95 / 86 / 93 ms)
                want = -(-(2 * n_decoders * n_lanes) // (8 * wps)) * 8
                want = min(max(8, want), (n_sm // 16) * 8)
                want = int(os.environ.get('L3C_DEC_SMS', want))           # bring-up knob
                part = E.partition_streams2(dev, want, 3 * n_lanes, 4 * n_lanes, n_enc)
            lanes = []
            for i in range(n_lanes):
                ln = DecodeLane()
                if part is not None:
                    a, b = part[0], part[1]
                    ln.dec = a[3 * i:3 * i + 3]
                    ln.main = b[4 * i]
                    ln.bld = b[4 * i + 1:4 * i + 4]
                    ln.partitioned = True
                else:
                    ln.dec = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(3)]
                    ln.bld = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(3)]
                    ln.main = torch.cuda.Stream(device=dev, priority=-1)
                    ln.partitioned = False
                lanes.append(ln)
            # two encode streams (default priority): the range-encoder launch of one batch is latency-bound (one
            # warp per stream, ~16 ms) and overlaps the convs of the next batch when they alternate
            encs = part[2] if part is not None else [torch.cuda.Stream(device=dev) for _ in range(n_enc)]
            cache[key] = (lanes, encs[0], encs)
        return cache[key]

    def encode_streams(self, dev, n_decoders, n_lanes=1):
        """The two default-priority streams for encodes that run beside decodes (alternate between them)."""
        return self.lanes(dev, n_decoders, n_lanes)[2]

    def _on_decoder_stream(self, lane, fn):
        """Run the launches of `fn` on the lane's first decoder stream (the decoders' SM group), ordered
        after what the current stream has queued and before what it queues next."""
        if lane is None or not lane.partitioned:
            return fn()
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(lane.dec[0]):
            lane.dec[0].wait_event(ev)
            out = fn()
            done = torch.cuda.Event()
            done.record(lane.dec[0])
        cur.wait_event(done)
        return out

    @_on_device
    def decode_device(self, blob, offs, lens, shapes, lane=None, tile=None):
        """Decode streams that already sit in HBM: `blob` uint8 device buffer (readable 4 bytes past
        every stream), offs/lens int64 [N][streams per image] in container order (coarse -> fine,
        channel-major), shapes [(C,H,W)] coarse -> fine.  Returns uint8 [N,3,H,W] on the device.
        `lane` (self.lanes()): the stream set of this decode when several decodes are in flight; the caller
        has made `lane.main` the current stream.  Without it the range decoders of the bottleneck scales run
        on the current stream and the RGB scale on lane 0 of a one-lane partition.
        `tile` = (th, tw): the streams are the tiles of a tiled container (ContainerLayout): thousands of short
        streams, so the three RGB channels are simply decoded one after the other, all tiles at once."""
        dev = blob.device
        N = offs.shape[0]
        if len(shapes) != self.net.scales + 1:
            raise ValueError('container has %d scales, model expects %d' % (len(shapes), self.net.scales + 1))
        K = self.net.config_ms.prob.K
        self._check_shapes(shapes)
        bn8, F_prev, S = None, None, None
        j0 = 0
        self._mark('start')
        if tile is not None:
            tile = (int(tile[0]), int(tile[1]))
        for idx, (scale, dmll, uniform) in enumerate(self.iter_scale_dmll()):
            C, H, W = shapes[idx]
            HW = H * W
            S = torch.empty(N, C, H, W, dtype=torch.uint8, device=dev)      # symbols in STREAM order (tile order if tiled)
            grid = tile_grid(H, W, tile)
            T = len(grid)
            g_off = np.array([o for (o, _) in grid], np.int64)[None, None, :]
            g_n = np.array([n for (_, n) in grid], np.int64)[None, None, :]
            desc = np.zeros(N * C * T, dtype=_lib.DEC_STREAM_DTYPE)
            d = desc.reshape(N, C, T)
            d['in'][:] = (blob.data_ptr() + offs[:, j0:j0 + C * T]).reshape(N, C, T)
            d['in_len'][:] = lens[:, j0:j0 + C * T].reshape(N, C, T)
            j0 += C * T
            d['n_sym'][:] = g_n
            plane = (np.arange(N)[:, None, None] * C + np.arange(C)[None, :, None]) * HW + g_off     # first symbol of the stream
            d['sym_out'][:] = S.data_ptr() + plane
            pitch = E.table_pitch(dmll.L)
            if uniform:
                row_dev = self._uniform(dmll.L, dev)[1]
                d['table'][:] = row_dev.data_ptr()
                d['row_pitch'][:] = 0
                self._on_decoder_stream(lane, lambda: E.ac_decode_streams(desc, dev, dmll.L))
            else:
                l, F_prev = self.net.get_P_nhwc(scale, bn8, F_prev if self.net._fuse_feat else None)
                self._mark('net%d' % scale)
                if tuple(l.shape[1:3]) != (H, W) or non_shared_get_K(l.shape[-1], C) != K:
                    raise ValueError('container scale %d (%dx%dx%d) does not match the model output %s'
                                     % (scale, C, H, W, tuple(l.shape)))
                table = torch.empty(N * C * HW * pitch, dtype=torch.int16, device=dev)
                d['table'][:] = table.data_ptr() + plane * (pitch * 2)
                d['row_pitch'][:] = pitch
                tg = dmll.targets(dev)
                if tile is not None:
                    if dmll.rgb_scale:
                        # channel c's means depend on the decoded channels < c of the same pixel: R, G, B one
                        # after the other, every tile of every image at once (N*T streams per launch)
                        for c in range(C):
                            E.dmll_build_table_tiled(l, S, tg, C, K, dmll.L, True, c, table, tile)
                            dc = np.ascontiguousarray(d[:, c, :]).reshape(-1)
                            self._on_decoder_stream(lane, lambda dc=dc: E.ac_decode_streams(dc, dev, dmll.L))
                    else:
                        E.dmll_build_table_tiled(l, S, tg, C, K, dmll.L, False, -1, table, tile)
                        self._on_decoder_stream(lane, lambda: E.ac_decode_streams(desc, dev, dmll.L))
                elif dmll.rgb_scale:
                    self._decode_rgb_pipelined(l, S, tg, C, K, dmll.L, table, d.reshape(N, C), dev, lane=lane)
                else:
                    E.dmll_build_table(l, S, tg, C, K, dmll.L, False, -1, table)
                    self._on_decoder_stream(lane, lambda: E.ac_decode_streams(desc, dev, dmll.L))
            if tile is not None:
                S = E.reorder_tiles(S, H, W, tile, False)                   # back to raster planes
            self._mark('uniform' if uniform else ('rgb' if dmll.rgb_scale and scale == 0 else 'S%d' % scale))
            if scale > 0:
                bn8 = E.symbols_to_values(S, self._symbol_values(scale, dmll, dev), self._rgb_shift(dev))
        return S

    # stage timing of a decode (bench.py / tools): set `codec.stage_events = []` before decode_device,
    # synchronise, then read stage_ms()
    stage_events = None

    def _mark(self, name):
        if self.stage_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.stage_events.append((name, ev))

    def stage_ms(self):
        """{stage: ms} of the decode recorded since `stage_events = []` (call after a synchronize)."""
        evs, self.stage_events = self.stage_events or [], None
        out = {}
        for (_, a), (name, b) in zip(evs[:-1], evs[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    def _check_shapes(self, shapes):
        """C, H, W of every scale come straight from the (untrusted) file, while the kernels index with
        the shapes the MODEL produces (each scale twice the size of the coarser one, C = 3 for the image
        and the RGB baselines, q.C for bottlenecks).  A crafted, truncated or wrong-config container must
        fail here, before any buffer is sized from it (the reference fails with a torch shape error)."""
        cfg = self.net.config_ms
        n = len(shapes)
        for idx, (C, H, W) in enumerate(shapes):
            scale = n - 1 - idx
            want_C = 3 if (scale == 0 or self.net._rgb) else cfg.q.C
            if C != want_C:
                raise ValueError('container scale %d has %d channels, model expects %d' % (scale, C, want_C))
            if H < 1 or W < 1:
                raise ValueError('container scale %d has an empty shape %dx%d' % (scale, H, W))
            if idx > 0:
                Hc, Wc = shapes[idx - 1][1:]
                if (H, W) != (2 * Hc, 2 * Wc):
                    raise ValueError('container scale %d is %dx%d, expected twice the coarser scale (%dx%d)'
                                     % (scale, H, W, 2 * Hc, 2 * Wc))

    def _symbol_values(self, scale, dmll, dev):
        """Value of every symbol of bottleneck `scale`: the `levels` LUT of the quantiser that produced
        it (nets[scale-1].enc.levels, a checkpoint parameter: net.py:123-127) so that the decoder feeds
        its nets exactly what the encoder fed (SURVEY finding 1); plain 0..255 for the RGB baselines."""
        if self.net._rgb:
            return dmll.values(dev)
        return self.net.nets[scale - 1].enc.levels.detach().float().contiguous()
