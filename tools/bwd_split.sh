#!/bin/bash
# Timing experiment: in-step time of the backward with individual kernels skipped (YN_BWD_DEBUG bit mask:
# 1 dgrad, 2 wgrad (+ intermediate-layer product), 4 heads, 8 direction, 16 heads kernel without its one-tile-ahead
# prefetch); gradients are wrong, only the times mean something.  Run every variant in ONE gpurun call: boxes differ by 3-4 %.
for m in 0 1 2 4 8 14; do
  YN_BWD_DEBUG=$m YANERF_TRAIN_GRAPH=0 timeout 60 python bench.py --workload train --steps 30 --warmup 5 2>/dev/null > /tmp/bs.json
  python - "$m" <<'PY'
import json, sys
d = json.load(open("/tmp/bs.json"))
print("skip", sys.argv[1], "step_ms", d["ms_per_step"], "bwd_ms", d["roofline"]["kernel_ms_per_step"]["yn_mlp_bwd"])
PY
done
