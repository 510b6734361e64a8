#!/bin/bash
# quick timing of the main kernels: bash scratch/quick.sh [steps] [extra bench.py flags]
s=${1:-5}; shift
python bench.py --steps $s --warmup 3 --no-cpu-baseline --no-e2e "$@" 2> gpurun_out/quick.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels_ms_per_launch']
print('value %.1f  enc %.1f dec %.1f  ms/step %.3f' % (d['value'], d['encode_gbs'], d['decode_gbs'], d['ms_per_step']), d.get('host_us_per_step',{}).get('trees'), d.get('host_us_per_step',{}).get('dectable'), d['gpu_launches'])
print({a: round(b,4) for a,b in k.items()})" || tail -5 gpurun_out/quick.err
