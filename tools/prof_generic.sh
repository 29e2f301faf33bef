#!/bin/bash
# ncu captures of the generic mixed-radix tile kernel (run on the GPU box via gpurun): 256/64 and 400/100
mkdir -p gpurun_out
for cfg in "256 64 256 128000" "400 100 64 100000"; do
  set -- $cfg
  python tools/prof_gl.py --n_fft $1 --hop $2 --batch $3 --samples $4 --iters 20
  ncu --set full --clock-control none --import-source on -k regex:mr_tile_kernel -s 3 -c 1 -f -o gpurun_out/r02_generic_mr_n$1 \
      python tools/prof_gl.py --n_fft $1 --hop $2 --batch $3 --samples $4 --iters 2 > gpurun_out/r02_generic_mr_n$1.log 2>&1
done
