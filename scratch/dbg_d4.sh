#!/bin/bash
for c in 1 2; do for d in 0 1; do
  echo "ctas $c dbg $d"; MH_DEC_WRITE_CTAS=$c MH_DEC_DBG=$d bash scratch/quick.sh 3 2>&1 | tail -1 | grep -o "'dec_write_kernel': [0-9.]*"
done; done
