#!/bin/bash
# A/B of two builds of libclair_b200.so on the same box: alternate runs of the headline bench (kernel-only numbers)
for round in 1 2; do
  for lib in old new; do
    if [ $lib = old ]; then export CLAIR_B200_LIB=$PWD/scratch/bin/libclair_old.so; else unset CLAIR_B200_LIB; fi
    for wl in c4 c1; do
      steps=200; [ $wl = c1 ] && steps=2000
      python bench.py --workload $wl --steps $steps --warmup 20 --no-extras --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', '$wl', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
    done
  done
done
