"""Bitcoding: encode an image to a `.l3c` file and decode it back -- the reference's API
(/root/reference/src/bitcoding/bitcoding.py:39-161), byte-compatible container, on top of the
batched sm_100a pipeline of codec.py.

Kept from the reference: `Bitcoding(blueprint, times, compare_with_theory)`, `encode(img, pout) ->
bpsp` (int64 CHW/1CHW image; asserts `pout` does not exist; crops -> `pout.part{i}`; bpsp is
file bytes * 8 / PADDED sub-pixels, crops combined weighted by unpadded H*W, :63-71,108-110),
`decode(pin) -> int64 1CHW`, the helper functions for the header fields and the magic separator.

Added (B200-first): `encode_batch` / `decode_batch` for many equally sized images per call.
"""
import os

import numpy as np
import torch

from . import auto_crop, pad, part_suffix_helper
from .blueprint import MultiscaleBlueprint
from .codec import BatchCodec, MAGIC
from .times import NoOp

_MAGIC_VALUE_SEP = MAGIC


class Bitcoding(object):
    def __init__(self, blueprint, times=NoOp, compare_with_theory=False, tile=None):
        """tile = (th, tw): write TILED containers (codec.ContainerLayout): every channel plane of every scale is
        cut into tiles coded as independent streams -- thousands of streams per batch instead of 18 per image,
        so encode and decode are no longer bound by one warp's serial chain.  Not readable by the reference and a
        few bytes per tile larger; decode() / decode_batch() recognise either layout.  Default: the reference's
        byte-compatible `.l3c` layout."""
        self.blueprint = blueprint
        self.compare_with_theory = compare_with_theory
        self.times = times
        self.tile = tile
        self.codec = BatchCodec(blueprint)

    # ------------------------------------------------------------------------------------------
    def _device(self):
        return self.blueprint.device

    def _prepare(self, img):
        """-> (uint8 [1,3,H',W'] padded, on device; padding tuple)."""
        factor = 2 ** self.blueprint.net.config_ms.num_scales
        _, _, H, W = img.shape
        padding_tuple = (0, 0, 0, 0)
        if H % factor != 0 or W % factor != 0:
            print(f'*** INFO: image shape ({H}X{W}) not divisible by {factor}, will pad.')
            img, padding_tuple = pad.pad(img, fac=factor, mode=MultiscaleBlueprint.get_padding_mode())
        return img.to(torch.uint8).to(self._device()).contiguous(), padding_tuple

    def encode(self, img, pout):
        """img: int64 (or uint8) CHW / 1CHW tensor with values 0..255; writes `pout`; returns bpsp."""
        assert not os.path.isfile(pout)
        if len(img.shape) == 3:
            img = img.unsqueeze(0)
        assert len(img.shape) == 4 and img.shape[0] == 1 and img.shape[1] == 3, img.shape
        assert img.dtype in (torch.int64, torch.uint8), img.dtype

        if auto_crop.needs_crop(img):
            print('Need to encode individual crops!')
            combinator = auto_crop.CropLossCombinator()
            crops = list(auto_crop.iter_crops(img))
            pouts = [pout + part_suffix_helper.make_part_suffix(i) for i in range(len(crops))]
            # the reference recurses into encode() per part and asserts on every part path
            # (bitcoding.py:57,63-71); stale higher-numbered parts of an earlier encode would be picked
            # up by decode() and stitched into a wrong image
            stale = [q for q in part_suffix_helper.existing_parts(pout) if q not in pouts]
            assert not any(os.path.isfile(q) for q in pouts) and not stale, \
                'part files of {} already exist: {}'.format(pout, [q for q in pouts if os.path.isfile(q)] + stale)
            bpsps = self._encode_many(crops, pouts)
            for crop, bpsp in zip(crops, bpsps):
                combinator.add(bpsp, np.prod(crop.shape[-2:]))
            return combinator.get_bpsp()
        return self._encode_many([img], [pout])[0]

    def _encode_many(self, imgs, pouts):
        """Images of equal (padded) shape are coded as one batch."""
        prepared = [self._prepare(im) for im in imgs]
        groups = {}
        for i, (t, pt) in enumerate(prepared):
            groups.setdefault((tuple(t.shape), pt), []).append(i)
        bpsps = [None] * len(imgs)
        for (shape, pt), idxs in groups.items():
            batch = torch.cat([prepared[i][0] for i in idxs], 0)
            with self.times.run('[-] encode forwardpass'):
                # without --compare_theory nobody reads the parameter tensors: the probability heads emit the
                # coding intervals directly (f16 mode; MultiscaleNetwork.forward(intervals_of=...))
                out = self.blueprint.forward(batch) if self.compare_with_theory else \
                    self.blueprint.forward_for_coding(batch)
            if self.compare_with_theory:
                with self.times.run('[-] get loss'):
                    loss_out = self.blueprint.get_loss(out)
            with self.times.run('[-] entropy coding'):
                datas, info = self.codec.encode_batch(batch, pt, out=out, tile=self.tile)
            num_subpixels = int(np.prod(shape))
            for k, i in enumerate(idxs):
                with open(pouts[i], 'wb') as f:
                    f.write(datas[k])
                bpsps[i] = len(datas[k]) * 8 / num_subpixels
            if self.compare_with_theory:
                total = sum(len(d) for d in datas) * 8 / (num_subpixels * len(idxs))
                theory = loss_out.nonrecursive_bpsps
                print('Bitrates:\ntheory:  {} => {:.3f}\nactual:  => {:.3f}'.format(
                    ' | '.join('{:.3f}'.format(b) for b in theory), sum(theory), total))
        return bpsps

    def decode(self, pin, _recurse_part=True):
        """-> decoded image, int64 1CHW (on the GPU)."""
        if _recurse_part and part_suffix_helper.contains_part_suffix(pin):
            paths = part_suffix_helper.iter_part_suffixes(pin)
            parts = self._decode_many(paths)
            print(f'Stitching {len(parts)} parts...')
            return auto_crop.stitch(parts)
        return self._decode_many([pin])[0]

    def _decode_many(self, paths):
        datas = []
        for p in paths:
            with open(p, 'rb') as f:
                datas.append(f.read())
        return self.decode_batch(datas)

    # ------------------------------------------------------------------------------------------
    # B200-first batch API
    def encode_batch(self, imgs_u8):
        """imgs_u8: uint8 [N,3,H,W] (any device) -> (list of container bytes, bpsp list)."""
        return self.encode_batch_begin(imgs_u8).finish()

    def encode_batch_begin(self, imgs_u8, stream=None):
        """Asynchronous encode_batch: enqueues the upload and the whole GPU side on `stream` (default:
        the current stream) and returns a job at once; job.finish() -> (containers, bpsps).  With
        `stream=self.side_stream(N)` the encode runs beside a decode_batch of another batch: the
        decode is bound by the serial range decoder and leaves most of the GPU idle."""
        factor = 2 ** self.blueprint.net.config_ms.num_scales
        pt = pad.padding_tuple(imgs_u8.shape[-2], imgs_u8.shape[-1], factor)
        stream = stream if stream is not None else torch.cuda.current_stream(self._device())
        with torch.cuda.stream(stream):
            x = imgs_u8.to(self._device(), non_blocking=True)
            if any(pt):
                x = torch.nn.functional.pad(x, pt, 'constant')
            x = x.contiguous()
            job = self.codec.encode_begin(x, pt, tile=self.tile)
        return _BatchEncodeJob(job, int(np.prod(x.shape[1:])))

    def side_stream(self, n_images, n_lanes=1):
        """Stream for encode_batch_begin() calls that should overlap decode_batch calls of `n_images` each."""
        return self.codec.encode_stream(self._device(), 3 * n_images, n_lanes)

    def decode_lanes(self, n_images, n_lanes=2):
        """Stream sets for `n_lanes` decode_batch calls in flight at the same time (codec.BatchCodec.lanes):
        `with torch.cuda.stream(lane.main): dec = bc.decode_batch(datas, lane=lane)` returns at once; the
        decode of one batch is bound by the serial chain of its range coder and leaves most of the GPU idle."""
        return self.codec.lanes(self._device(), 3 * n_images, n_lanes)[0]

    def decode_batch(self, datas, lane=None):
        """list of container bytes (any mix of shapes) -> list of int64 1CHW tensors (GPU)."""
        from .codec import parse_container
        keys = {}
        for i, d in enumerate(datas):
            pt, scales = parse_container(d)
            keys.setdefault((pt, tuple((C, H, W) for (C, H, W, _) in scales)), []).append(i)
        outs = [None] * len(datas)
        for (pt, _), idxs in keys.items():
            S, _ = self.codec.decode_batch([datas[i] for i in idxs], to_host=False, lane=lane)
            for k, i in enumerate(idxs):
                img = S[k:k + 1].long()
                if any(pt):
                    img = pad.undo_pad(img, *pt)
                outs[i] = img
        return outs


class _BatchEncodeJob:
    def __init__(self, job, nsub):
        self.job, self.nsub = job, nsub

    def finish(self):
        datas, _ = self.job.finish(to_host=True)
        return datas, [len(d) * 8 / self.nsub for d in datas]


# --- header field helpers (bitcoding.py:326-375), little-endian -------------------------------
def write_bytes(f, ts, xs):
    for t, x in zip(ts, xs):
        f.write(t(x).tobytes())


def read_bytes(f, ts):
    return [np.frombuffer(f.read(t().itemsize), t, count=1)[0] for t in ts]


def write_shape(shape, fout):
    assert len(shape) == 4 and shape[0] == 1, shape
    C, H, W = shape[1:]
    assert C < 2 ** 8 and H < 2 ** 16 and W < 2 ** 16, shape
    write_bytes(fout, [np.uint8, np.uint16, np.uint16], (C, H, W))
    return 5


def read_shapes(fin):
    return tuple(map(int, read_bytes(fin, [np.uint8, np.uint16, np.uint16])))


def write_num_bytes_encoded(num_bytes, fout):
    assert num_bytes < 2 ** 32
    write_bytes(fout, [np.uint32], [num_bytes])
    return 2


def read_num_bytes_encoded(fin):
    return int(read_bytes(fin, [np.uint32])[0])


def write_padding_tuple(padding_tuple, fout):
    assert len(padding_tuple) == 4
    write_bytes(fout, [np.uint16] * 4, padding_tuple)


def read_padding_tuple(fin):
    return tuple(map(int, read_bytes(fin, [np.uint16] * 4)))
