"""Pins ac_oracle.c against the reference's own torchac.cpp (compiled unmodified into
oracle/_ref by build_ref.py): KATs of SURVEY.md section 8c + random tables, byte-for-byte.
Run here (needs /root/reference or a prebuilt oracle/_ref).  TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ac, build_ref  # noqa: E402


def main():
    ref = build_ref.load()
    if ref is None:
        print('reference backend unavailable; nothing pinned')
        return 1
    rng = np.random.default_rng(7)
    n_cases = 0
    for Lp, n_sym in [(26, 1), (26, 8), (26, 4097), (257, 3), (257, 5000), (6, 777), (2, 64)]:
        for trial in range(6):
            L = Lp - 1
            # random strictly increasing rows with >= 1 count per symbol, first entry may be > 0
            w = rng.integers(1, 4000 if trial % 2 else 40, size=(n_sym, L)).astype(np.float64)
            w = w / w.sum(1, keepdims=True) * (65536 - Lp - 40)
            c = np.floor(np.cumsum(w, 1)).astype(np.int64) + np.arange(1, L + 1)
            lead = rng.integers(0, 30, size=(n_sym, 1))
            cdf = np.concatenate([lead, c[:, :-1] + lead, np.zeros((n_sym, 1), np.int64)], 1)
            cdf = cdf.astype(np.uint16)
            assert (np.diff(cdf[:, :-1].astype(np.int64), axis=1) > 0).all()
            sym = rng.integers(0, L, size=n_sym).astype(np.int16)
            t_cdf = torch.from_numpy(cdf.view(np.int16).copy()).reshape(1, 1, n_sym, Lp)
            want = ref.encode_cdf(t_cdf, torch.from_numpy(sym.copy()))
            got = ac.encode(cdf, sym)
            assert got == want, (Lp, n_sym, trial, len(got), len(want))
            dec_ref = ref.decode_cdf(t_cdf, want).numpy()
            dec = ac.decode(cdf, want)
            assert (dec == dec_ref).all() and (dec == sym).all()
            # truncated / garbage input: decoder zero-fills, must still agree with the reference
            junk = bytes(rng.integers(0, 256, size=max(1, len(want) // 2)).astype(np.uint8))
            assert (ac.decode(cdf, junk) == ref.decode_cdf(t_cdf, junk).numpy()).all()
            n_cases += 1
    # long underflow runs: every symbol straddles the midpoint, the owed ("pending") bits pile up far
    # beyond 32 and across many symbols before a release or the terminator flushes them
    # (torchac.cpp:196-206, 209-219) -- the path the GPU encoder's batched emission must reproduce
    n_under = 0
    for n_sym in (5, 33, 1000, 4099):
        for trial in range(3):
            half = 32768
            rows = np.zeros((n_sym, 4), np.int64)
            for i in range(n_sym):
                rows[i, :3] = [0, half - int(rng.integers(1, 200)), half + int(rng.integers(1, 200))]
            sym = np.ones(n_sym, np.int16)
            if trial < 2:
                for i in rng.integers(0, n_sym, size=max(1, n_sym // 37)):
                    sym[i] = int(rng.integers(0, 3))
            cdf = rows.astype(np.uint16)
            t_cdf = torch.from_numpy(cdf.view(np.int16).copy()).reshape(1, 1, n_sym, 4)
            want = ref.encode_cdf(t_cdf, torch.from_numpy(sym.copy()))
            assert ac.encode(cdf, sym) == want, ('underflow', n_sym, trial)
            assert (ac.decode(cdf, want) == ref.decode_cdf(t_cdf, want).numpy()).all()
            n_under += 1
    n_cases += n_under
    # KATs
    row25 = ac.uniform_cdf_row(25)
    assert row25.tolist()[:4] == [0, 2621, 5243, 7864] and row25[-1] == 0 and row25[-2] == 62915
    assert ac.encode(row25, np.array([0, 1, 2, 3, 24, 23, 12, 12], np.int16)).hex() == '0071e1d840'
    row256 = ac.uniform_cdf_row(256)
    assert ac.encode(row256, np.array([0, 255, 128, 1, 254, 77], np.int16)).hex() == '00ff8001fe4d40'
    assert ac.encode(row25, np.array([0], np.int16)).hex() == '04'
    assert ac.encode(row25, np.array([24], np.int16)).hex() == 'f8'
    print('pinned: %d random cases (incl. %d long-underflow ones) + KAT1/1b/3 byte-identical to oracle/_ref'
          % (n_cases, n_under))
    return 0


if __name__ == '__main__':
    sys.exit(main())
