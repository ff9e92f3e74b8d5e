for l in 1 2 3; do python bench.py --workload rgb_shared --value-only --lanes $l 2>&1 | tail -1 | cut -c60-200; done
L3C_DEC_WARPS_PER_SM=16 python bench.py --workload rgb_shared --value-only --lanes 2 2>&1 | tail -1 | cut -c60-200
python bench.py --workload crops --no-comparators --no-cpu-baseline > gpurun_out/bench_crops2.json 2> gpurun_out/bench_crops2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_crops2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['sequential'], d['breakdown']['decode_ms'], d['config']['workload'][:80])
PY
