#!/bin/bash
# A/B of builds on the same box (box-to-box variance is 1-2 %): bash scratch/ab.sh [-c config] scratch/lib_a.so scratch/lib_b.so ...
# prints value, ms per step and the main kernels' times for every build, two alternating rounds
cfg=markov
if [ "$1" = "-c" ]; then cfg=$2; shift 2; fi
for r in 1 2; do for l in "$@"; do cp $l markov-huffman-coding_b200/libmh_gpu.so
  python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels_ms_per_launch']
print('$l', 'value %.1f ms/step %.3f' % (d['value'], d['ms_per_step']), ' '.join('%s %.4f' % (a.replace('_kernel',''), b) for a,b in k.items() if b > 0.05))"
done; done
