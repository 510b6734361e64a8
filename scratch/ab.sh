#!/bin/bash
# A/B of two builds on the same box: bash scratch/ab.sh scratch/lib_prev.so scratch/lib_new.so
for r in 1 2; do for l in "$@"; do cp $l markov-huffman-coding_b200/libmh_gpu.so; echo "$l: $(bash scratch/quick.sh 10 | tail -1 | grep -o "'encode_kernel': [0-9.]*\|'dec_sync_kernel': [0-9.]*\|'dec_write_kernel': [0-9.]*" | tr '\n' ' ')"; done; done
