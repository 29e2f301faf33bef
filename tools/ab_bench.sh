#!/bin/bash
# A/B of two library builds on the headline bench (sustained, power-capped): tools/ab_bench.sh <other.so> [reps]
other=$1; reps=${2:-2}
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; print("value %.3f M  ms/step %.2f  sustained %.4f ms (%.3f)  burst %.4f ms (%.3f)  sm %s MHz" % (d["value"]/1e6, d["ms_per_step"], r["ms_per_launch"], r["frac"], r["burst_ms_per_launch"], r["burst_frac"], d["clocks"]["sm_mhz"]))'
for rep in $(seq $reps); do
  echo -n "current : "; python bench.py --steps 10 --warmup 3 --configs none --no-cpu 2>/dev/null | python -c "$pick"
  echo -n "other   : "; SPECINV_B200_LIB=$other python bench.py --steps 10 --warmup 3 --configs none --no-cpu 2>/dev/null | python -c "$pick"
done
