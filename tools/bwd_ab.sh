#!/bin/bash
# heads-kernel time = backward with everything minus backward without it (YN_BWD_DEBUG=4), same box, graph-mode bench
for m in 0 4 0 4; do
  YN_BWD_DEBUG=$m timeout 60 python bench.py --workload train --steps 50 --warmup 5 2>/dev/null > /tmp/bs.json
  python - "$m" <<'PY'
import json, sys
d = json.load(open("/tmp/bs.json"))
print("debug", sys.argv[1], "step_ms", d["ms_per_step"], "bwd_ms", d["roofline"]["kernel_ms_per_step"]["yn_mlp_bwd"])
PY
done
