#!/bin/bash
# A/B of two builds of the library on the same box, interleaved: put them at tools/_ab/libold.so and tools/_ab/libnew.so
# (git-ignored), then  gpurun -- bash tools/ab_bench.sh ; rebuild the library afterwards (the script overwrites it).
for round in 1 2 3; do
  for v in old new; do
    cp tools/_ab/lib$v.so yet-another-nerf_b200/libyanerf_b200.so
    python bench.py --only none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['clocks']['sm_mhz'])"
  done
done
