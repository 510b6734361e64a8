#!/bin/bash
# quick timing of the four main kernels: bash scratch/quick.sh [steps]
python bench.py --steps ${1:-5} --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels_ms_per_launch']
print('value %.1f  enc %.1f dec %.1f  ms/step %.3f' % (d['value'], d['encode_gbs'], d['decode_gbs'], d['ms_per_step']), d.get('host_us_per_step'))
print({a: round(b,4) for a,b in k.items()})"
