#!/bin/bash
# A/B timing of two builds of the library on the BASELINE configs: tools/ab_configs.sh <other.so> [configs...]
other=$1; shift
cfgs=${@:-cfg2 cfg4 cfg5 cfg1}
for rep in 1 2; do
  echo "== current build (rep $rep)"; python tools/bench_configs.py $cfgs
  echo "== $other (rep $rep)"; SPECINV_B200_LIB=$other python tools/bench_configs.py $cfgs
done
