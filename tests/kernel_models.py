"""Plain-Python models of the two range-coder ideas the CUDA kernels rest on (TEST INFRASTRUCTURE):
the (low, r) coder state with the collapsed renormalisation, and the encoder's batched, order-free bit
emission.  They let the CPU suite check the algorithms against the oracle (and through it the
reference, torchac.cpp:152-227) without a GPU; the kernels themselves are checked in the -m gpu suite."""
import numpy as np

M32 = 0xFFFFFFFF


def clz32(x):
    return 32 - int(x).bit_length()


def shl(x, n):
    return (x << n) & M32 if n < 32 else 0


def coder_records(intervals):
    """intervals: iterable of (c_lo, c_hi) in 1/65536 units.  Returns per symbol (low before the
    shift, k, u) and the final low -- the quantities ac_encode_kernel keeps, computed with its state
    (low, r = high - low):  width = hi16 - lo16, low' = (low << s) & 0x7FFFFFFF, r' = (width << s) - 1."""
    low, r = 0, M32
    recs = []
    for c_lo, c_hi in intervals:
        hi16 = ((r * c_hi + c_hi) >> 16) & M32
        lo16 = ((r * c_lo + c_lo) >> 16) & M32
        width = (hi16 - lo16) & M32
        low = (low + lo16) & M32
        high = (low + width - 1) & M32
        k = clz32(low ^ high)
        y = low & ~high & M32
        u = clz32(~shl(shl(y, k), 1) & M32)
        recs.append((low, k, u))
        low = shl(shl(low, k), u) & 0x7FFFFFFF
        r = (shl(shl(width, k), u) - 1) & M32
    return recs, low


def emit_batched(recs, final_low, batch=32):
    """Bit strings of `batch` symbols at a time, each computed independently from prefix sums:
    string(j) = k>0 ? b0, pending(j) x !b0, remaining k-1 bits : nothing; pending(j) = owed underflow
    bits when symbol j is reached (sum of u since the last symbol with k>0).  Then the terminator
    (torchac.cpp:209-219) as a one-bit record met with pending+1 owed bits."""
    bits = []
    pending = 0

    def run(batch_recs, pending):
        k = np.array([r[1] for r in batch_recs], np.int64)
        u = np.array([r[2] for r in batch_recs], np.int64)
        excl_u = np.concatenate([[0], np.cumsum(u)[:-1]])
        flagged = k > 0
        out = [None] * len(batch_recs)
        for j, (low, kj, uj) in enumerate(batch_recs):           # every j only reads the prefix sums
            before = np.nonzero(flagged[:j])[0]
            pend = (excl_u[j] - excl_u[before[-1]]) if len(before) else pending + excl_u[j]
            if kj == 0:
                out[j] = ''
                continue
            top = format(low >> (32 - kj), '0%db' % kj)
            out[j] = top[0] + ('1' if top[0] == '0' else '0') * int(pend) + top[1:]
        last = np.nonzero(flagged)[0]
        new_pending = (np.sum(u) - excl_u[last[-1]]) if len(last) else pending + np.sum(u)
        return ''.join(out), int(new_pending)

    for b0 in range(0, len(recs), batch):
        s, pending = run(recs[b0:b0 + batch], pending)
        bits.append(s)
    fin = 0 if final_low < 0x40000000 else 1
    s, _ = run([(fin << 31, 1, 0)], pending + 1)
    bits.append(s)
    allbits = ''.join(bits)
    allbits += '0' * (-len(allbits) % 8)
    return bytes(int(allbits[i:i + 8], 2) for i in range(0, len(allbits), 8))


def decode_model(cdf, data, n_sym):
    """Decoder with the kernels' state (low, r, dv = value - low): per symbol the passing entries are
    those with mulhi(cdf[m] << 16, r + 1) <= dv (no division; equals cdf[m] <= count of
    torchac.cpp:327), the update is width/low as in coder_records and dv' = ((dv - lo16) << s) | next s
    code bits.  A code value outside [low, high] takes the reference's modular count (the kernels' replay path)."""
    bits = ''.join(format(b, '08b') for b in data)
    pos = [0]

    def take(n):
        s = bits[pos[0]:pos[0] + n]
        pos[0] += n
        s = s + '0' * (n - len(s))                              # zero fill past the end (torchac.cpp:109-112)
        return int(s, 2) if n else 0

    low, r, dv = 0, M32, take(32)
    out = []
    L = cdf.shape[1] - 1
    for i in range(n_sym):
        span = r + 1
        if dv <= r:
            passing = [m for m in range(L) if ((int(cdf[i, m]) << 16) * span) >> 32 <= dv]
        else:
            # code value outside [low, high] (garbage / foreign input): the reference's count in its
            # modular 64-bit arithmetic (torchac.cpp:327), as the kernels' replay path computes it
            off = ((low + dv) & M32) - low + 1
            count = ((((off & 0xFFFFFFFFFFFFFFFF) * 65536 - 1) & 0xFFFFFFFFFFFFFFFF) // span) & 0xFFFF
            passing = [m for m in range(L) if int(cdf[i, m]) <= count]
        m = passing[-1] if passing else 0
        out.append(m)
        if i == n_sym - 1:
            break
        c_lo = int(cdf[i, m])
        c_hi = 65536 if m == L - 1 else int(cdf[i, m + 1])
        hi16 = ((r * c_hi + c_hi) >> 16) & M32
        lo16 = ((r * c_lo + c_lo) >> 16) & M32
        width = (hi16 - lo16) & M32
        lo = (low + lo16) & M32
        hi = (lo + width - 1) & M32
        d = (dv - lo16) & M32
        k = clz32(lo ^ hi)
        u = clz32(~shl(shl(lo & ~hi & M32, k), 1) & M32)
        s = k + u
        low = shl(lo, s) & 0x7FFFFFFF
        r = (shl(width, s) - 1) & M32
        dv = (shl(d, s) | take(s)) & M32
    return np.array(out, np.int16)
