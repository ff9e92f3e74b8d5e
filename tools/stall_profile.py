"""Per-instruction stall profile from an `ncu --page source --csv` export: walks the hottest loop
in address order and prints samples per instruction (samples ~ cycles the warp sat on it)."""
import csv, sys
path = sys.argv[1]; lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0; hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
base = int(data[0][0], 16)
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot)
for r in data:
    off = int(r[0], 16) - base
    if not (lo <= off < hi): continue
    n = int(r[ix['# Samples']])
    top = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print('%05x %6d %8s  %-60s %s' % (off, n, r[ix['Instructions Executed']], r[ix['Source']].strip()[:60],
                                      ' '.join('%s:%d' % (h, c) for c, h in top if c)))
