#!/bin/bash
# needs the instrumented library: make -C yet-another-nerf_b200/csrc clean all INSTRUMENT=1
# Timing experiments on the training step (gradients are wrong under these flags, only the times mean something).
#   YN_BWD_DEBUG bit mask: 1 skip dgrad, 2 skip wgrad (+ intermediate-layer product), 4 skip heads, 8 skip direction,
#                          32 dgrad stores into an L2-resident 64-tile window, 64 wgrad reads from such a window
#   YN_FWD_DEBUG bit 4:    forward stash stores into an L2-resident window
# Run every variant in ONE gpurun call: boxes differ by 3-4 %.
for cfg in "0 0" "0 2" "0 34" "0 1" "0 65" "4 0" "0 14"; do
  set -- $cfg
  YN_FWD_DEBUG=$1 YN_BWD_DEBUG=$2 YANERF_TRAIN_GRAPH=0 timeout 120 python bench.py --only train --steps 2 --train-steps 30 2>/dev/null > /tmp/bs.json
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open("/tmp/bs.json"))["train"]
k = d["roofline"]["kernel_ms_per_step"]
print("fwd_debug", sys.argv[1], "bwd_debug", sys.argv[2], "step_ms", d["ms_per_step"], "fwd_ms", k["yn_mlp_fwd"], "bwd_ms", k["yn_mlp_bwd"])
PY
done
