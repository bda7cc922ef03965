#!/bin/bash
# A/B runs of the device-resident step under environment variants: one bench line per variant, reduced to the numbers
# that matter while tuning (ms per step, per-phase kernel ms, e2e ms).
#   bash tools/ab.sh TAG "NAME1:ENV1=V ENV2=V" "NAME2:..."
set +e
O=gpurun_out; TAG=$1; shift; mkdir -p $O
for spec in "default:" "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_ab_${name}.json 2> $O/${TAG}_ab_${name}.err
  python - $O/${TAG}_ab_${name}.json "$name" "$envs" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    k = d["kernel_ms_per_step"]
    print("%-14s step %.3f ms (p1 %.3f, p2 %.3f, merge+tally %.3f) e2e %.2f ms  frac %.3f  parity %s  [%s]" % (
        sys.argv[2], d["ms_per_step"], k["sort1+poa_dp1_kernel"], k["sort2+poa_dp2_kernel"], k["merge+tally"], d["e2e"]["ms_per_step"],
        d["roofline"]["frac"], d["config"]["parity"][:9], sys.argv[3]))
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace(".json", ".err")).read()[-600:])
PY
done
