"""Micro-benchmark of the range-coder kernels alone: ns per symbol per stream."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from l3c_pytorch_b200 import engine as E, _lib

def run(L, n_streams, n_sym, peaked, rare=False, own_tables=False, uniform=False):
    dev = torch.device('cuda')
    rng = np.random.default_rng(0)
    pitch = E.table_pitch(L)
    # one random table shared by all streams (content is irrelevant for timing, sizes are not)
    w = rng.integers(1, 40, size=(n_sym, L)).astype(np.float64)
    if peaked:
        w[np.arange(n_sym), rng.integers(0, L, n_sym)] *= (2000 if not rare else 3e7)
    w = w / w.sum(1, keepdims=True) * (65536 - L - 40)
    c = np.floor(np.cumsum(w, 1)).astype(np.int64) + np.arange(1, L + 1)
    cdf = np.concatenate([np.zeros((n_sym, 1), np.int64), c[:, :-1]], 1)          # [n_sym, L]
    hi_all = np.concatenate([cdf[:, 1:], np.full((n_sym, 1), 65536)], 1)
    sym = np.array([rng.choice(L, p=(wi / wi.sum())) for wi in w[:2000]])
    sym = np.resize(sym, n_sym).astype(np.int64)
    if rare:      # symbols drawn uniformly under a peaked model: ~15 bits each (random-init L3C on noise)
        sym = rng.integers(0, L, n_sym).astype(np.int64)
    if uniform:       # the uniform-prior scale: ONE row (l3c_uniform_cdf_row) shared by every symbol, row pitch 0
        row = E.uniform_cdf_row(L).astype(np.int64)
        row[L] = 65536
        cdf = np.repeat(row[None, :L], n_sym, 0); hi_all = np.repeat(row[None, 1:], n_sym, 0)
        sym = rng.integers(0, L, n_sym).astype(np.int64)
    lo = cdf[np.arange(n_sym), sym]; hi = hi_all[np.arange(n_sym), sym]
    iv = torch.from_numpy((lo | ((hi - 1) << 16)).astype(np.uint32).view(np.int32)).to(dev)
    tab = np.zeros((n_sym, pitch), np.uint16); tab[:, :L] = cdf
    tab_dev = torch.from_numpy(tab.view(np.int16)).to(dev)
    cap = ((n_sym * 17 + 7) // 8 + 64 + 3) & ~3
    slots = torch.zeros(n_streams * cap, dtype=torch.uint8, device=dev)
    desc = np.zeros(n_streams, dtype=_lib.ENC_STREAM_DTYPE)
    desc['intervals'] = iv.data_ptr(); desc['n_sym'] = n_sym; desc['out_cap'] = cap
    desc['out'] = slots.data_ptr() + cap * np.arange(n_streams)
    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
        return best
    ddev = E._desc_to_device(desc, dev)
    lens = torch.empty(n_streams, dtype=torch.int32, device=dev)
    t_enc = timed(lambda: E.check(E.lib.l3c_ac_encode_streams(E._ptr(ddev), n_streams, E._ptr(lens), E._stream_ptr())))
    nbytes = int(lens[0])
    out = torch.zeros(n_streams * n_sym, dtype=torch.uint8, device=dev)
    dd = np.zeros(n_streams, dtype=_lib.DEC_STREAM_DTYPE)
    if own_tables:      # every stream streams its own rows from HBM (as in the real pipeline)
        big = tab_dev.repeat(n_streams, 1)
        dd['table'] = big.data_ptr() + np.arange(n_streams, dtype=np.int64) * (n_sym * pitch * 2)
    else:
        dd['table'] = tab_dev.data_ptr()
    dd['in'] = slots.data_ptr() + cap * np.arange(n_streams)
    dd['sym_out'] = out.data_ptr() + n_sym * np.arange(n_streams); dd['row_pitch'] = 0 if (uniform and not own_tables) else pitch
    dd['n_sym'] = n_sym; dd['in_len'] = nbytes
    dddev = E._desc_to_device(dd, dev)
    t_dec = timed(lambda: E.check(E.lib.l3c_ac_decode_streams(E._ptr(dddev), n_streams, L, 0, n_sym, E._stream_ptr())))
    ok = bool((out.reshape(n_streams, n_sym).cpu().numpy() == sym[None, :]).all())
    print(json.dumps(dict(L=L, streams=n_streams, n_sym=n_sym, peaked=peaked, rare=rare, own_tables=own_tables, uniform=uniform, bits_per_sym=nbytes * 8 / n_sym,
                          enc_ns_per_sym=t_enc * 1e6 / n_sym, dec_ns_per_sym=t_dec * 1e6 / n_sym, ok=ok)))

if __name__ == '__main__':
    only = int(sys.argv[sys.argv.index('--only') + 1]) if '--only' in sys.argv else None
    if '--custom' in sys.argv:        # --custom L,streams,n_sym,peaked,rare,own_tables,uniform  (0/1 flags)
        v = [int(x) for x in sys.argv[sys.argv.index('--custom') + 1].split(',')]
        run(v[0], v[1], v[2], bool(v[3]), bool(v[4]), bool(v[5]), bool(v[6]))
        sys.exit(0)
    for ci, (L, ns, n, pk, rare, own, uni) in enumerate([(256, 48, 65536, True, False, False, False), (256, 48, 65536, True, True, False, False),
                                      (25, 80, 65536, True, True, False, False), (256, 48, 65536, True, True, True, False),
                                      (25, 80, 65536, True, True, True, False), (256, 96, 16384, False, False, False, True),
                                      (25, 80, 4096, False, False, False, True)]):
        if only is None or only == ci:
            run(L, ns, n, pk, rare, own, uni)
