#!/bin/bash
# Perf experiment: time of the MMA stream alone (KG_FILTER_DEBUG=3: no expansion, no epilogue work) against UMMA N
# (p_pad = 16 ceil((P+1)/16)), A from tensor memory vs (bit 32) A from shared memory.  Results of these runs are WRONG
# by design; only the kernel time is read.
tag=${1:-probe}
mkdir -p gpurun_out
run() {  # dbg phenos a_words
  KG_FILTER_DEBUG=$1 KG_FILTER_A_WORDS=$3 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --kinship-rows 0 --e2e-buffers 1 \
      --phenos $2 > gpurun_out/${tag}_d$1_p$2_a$3.json 2> gpurun_out/${tag}_d$1_p$2_a$3.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_d$1_p$2_a$3.json").read())
    k = d["roofline"]["kernels"]
    ms = k["scan_filter"]["ms_per_launch"]; rows = k["scan_filter"]["rows_per_launch"]
    clk_blk = ms * 1e-3 * 1.965e9 / (rows / 128 / 148)
    print("dbg=$1 P=$2 a_words=$3: filter ms/launch %.4f  clk/block %.0f  clk/MMA %.1f  step ms %.4f" % (ms, clk_blk, clk_blk / 36, d["ms_per_step"]))
except Exception as e:
    print("dbg=$1 P=$2 failed", e)
PY
}
for p in 31 47 63 79 95 111 127; do run 3 $p 8; done
run 3 101 9
run 35 101 9
run 35 63 8
run 35 127 8
run 0 101 9
