#!/bin/bash
# sweep of D4 launch shapes: threads per CTA x CTAs per SM
for t in 256 384 512 768 1024; do for c in 1 2 3 4; do
  r=$(MH_DEC_WRITE_THREADS=$t MH_DEC_WRITE_CTAS=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels_ms_per_launch']; print(k['dec_sync_kernel'], k['dec_write_kernel'])")
  echo "threads $t ctas $c : $r"
done; done
