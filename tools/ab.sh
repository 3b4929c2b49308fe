#!/bin/bash
# usage: tools/ab.sh tag [ENV=VAL ...] -- runs bench.py (no CPU leg) and prints value + ms
tag=$1; shift
env NBIOT_BENCH_NO_CPU=1 "$@" python bench.py --steps 8 --warmup 3 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/ab_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], '%.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['config']['launch_geometry'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
