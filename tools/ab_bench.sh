#!/bin/bash
# A/B of two builds of the library on the same box, interleaved: put them at tools/_ab/libold.so and tools/_ab/libnew.so
# (git-ignored), then  gpurun -- bash tools/ab_bench.sh [bench.py flags, default "--only none"] ; rebuild the library
# afterwards (the script overwrites it).  Prints: build, render rays/s, ms per image, fine-pass ms, SM MHz[, train rays/s, ms].
FLAGS=${1:---only none}
for round in 1 2 3; do
  for v in old new; do
    cp tools/_ab/lib$v.so yet-another-nerf_b200/libyanerf_b200.so
    python bench.py $FLAGS --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d.get('train')
print('$v', d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['clocks']['sm_mhz'], *( [t['value'], t['ms_per_step']] if t else []))"
  done
done
