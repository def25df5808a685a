#!/bin/bash
# Perf experiment: which warp role bounds kg_scan_filter_kernel?  KG_FILTER_DEBUG bits: 1 skip expansion, 2 skip epilogue
# work, 4 skip MMAs, 8 skip bulk loads, 16 a_empty by plain arrive.  Results of these runs are WRONG by design; only the
# kernel time is read.   gpurun -- 'bash profiles/filter_dbg_sweep.sh <tag>'
tag=${1:-sweep}
mkdir -p gpurun_out
for dbg in 0 2 3 6 7 10 11 14; do
  KG_FILTER_DEBUG=$dbg timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --kinship-rows 0 --e2e-buffers 1 \
      > gpurun_out/${tag}_dbg${dbg}.json 2> gpurun_out/${tag}_dbg${dbg}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_dbg${dbg}.json").read())
    k = d["roofline"]["kernels"]
    print("dbg=${dbg}", "filter ms/launch", round(k["scan_filter"]["ms_per_launch"], 4), "refine", round(k.get("scan_refine", {}).get("ms_per_launch", 0), 4), "step ms", round(d["ms_per_step"], 4))
except Exception as e:
    print("dbg=${dbg} failed", e)
PY
done
