import os
print({k: v for k, v in os.environ.items() if 'INJ' in k.upper() or 'NSIGHT' in k.upper() or 'NV_' in k.upper() or 'LD_PRELOAD' in k})
print(sorted(set(l.split()[-1] for l in open('/proc/self/maps') if 'nsight' in l.lower() or 'inject' in l.lower()))[:8])
