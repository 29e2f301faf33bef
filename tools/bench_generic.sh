#!/bin/bash
# event timings of the generic (mixed-radix) kernels on a few shapes: python tools/prof_gl.py per shape
for cfg in "256 64 256 128000" "128 32 256 64000" "400 100 64 100000" "1000 250 64 240000" "1536 384 64 240000" "8192 2048 8 960000"; do
  set -- $cfg
  SPECINV_FORCE_GENERIC=1 python tools/prof_gl.py --n_fft $1 --hop $2 --batch $3 --samples $4 --iters 20 || true
done
SPECINV_FORCE_GENERIC=1 python tools/prof_gl.py --n_fft 1024 --hop 256 --batch 64 --samples 240000 --iters 20
SPECINV_FORCE_GENERIC=1 python tools/prof_gl.py --algo admm --n_fft 400 --hop 100 --batch 64 --samples 100000 --iters 20
