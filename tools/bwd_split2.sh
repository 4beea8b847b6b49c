#!/bin/bash
# needs the instrumented library: make -C yet-another-nerf_b200/csrc clean all INSTRUMENT=1
# Where do the training forward and the data-gradient kernel lose time?  (timing experiments, wrong results)
#   YN_FWD_DEBUG: 8 no sign masks, 16 no stash stores, 4 stash into an L2-resident window
#   YN_BWD_DEBUG: 14 = dgrad only; +128 no mask application, +256 no gradient-stash stores
for cfg in "0 14" "8 14" "16 14" "24 14" "0 142" "0 270" "0 398"; do
  set -- $cfg
  YN_FWD_DEBUG=$1 YN_BWD_DEBUG=$2 timeout 120 python bench.py --profile train --steps 30 2>/dev/null > /tmp/bs.json
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open("/tmp/bs.json"))
k = d["roofline"]["kernel_ms_per_step"]
print("fwd_debug", sys.argv[1], "bwd_debug", sys.argv[2], "fwd_ms", k["yn_mlp_fwd"], "bwd(dgrad only)_ms", k["yn_mlp_bwd"])
PY
done
