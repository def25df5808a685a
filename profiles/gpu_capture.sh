#!/bin/bash
# One gpurun call: launch list of a short job + ncu --set full captures of the scan's kernels (never a bench value).
#   gpurun --timeout 1500 -- 'bash profiles/gpu_capture.sh <tag>'
# Outputs land in gpurun_out/<tag>_*; summaries are copied by hand into profiles/.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> $out/${tag}_gpu.txt
B="python bench.py --steps 6 --warmup 3 --job-rows 600000000 --no-cpu-baseline --no-parity --kinship-rows 0 --e2e-buffers 1"

echo "== launch list (whole short job: cold phase + steady rounds)" ; date
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_launches_bench.log 2>&1; echo "rc=$?"

echo "== ncu full: filter (steady round), heap replay (cold round), pair kernel, kinship, narrow-table filter" ; date
# the warm-up scans 3 of the 6 steps (about 30 rounds), the timed job has about 40: launch 60 of a per-round kernel is a
# steady 2^24-row round of the timed job, launch 45 a late cold round
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:kg_scan_filter -s 60 -c 1 -f -o $out/${tag}_prof_filter $B > $out/${tag}_prof_filter.log 2>&1; echo "rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:kg_select_replay -s 45 -c 1 -f -o $out/${tag}_prof_replay $B > $out/${tag}_prof_replay.log 2>&1; echo "rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:kg_scan_pair -s 60 -c 1 -f -o $out/${tag}_prof_pair $B > $out/${tag}_prof_pair.log 2>&1; echo "rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:kg_kinship_tc -s 3 -c 1 -f -o $out/${tag}_prof_kinship \
    python bench.py --steps 2 --warmup 3 --job-rows 50000000 --no-cpu-baseline --no-parity --e2e-buffers 1 > $out/${tag}_prof_kinship.log 2>&1; echo "rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:kg_scan_filter -s 60 -c 1 -f -o $out/${tag}_prof_filter_n241 \
    python bench.py --samples 241 --steps 6 --warmup 3 --job-rows 1200000000 --no-cpu-baseline --no-parity --kinship-rows 0 --e2e-buffers 1 > $out/${tag}_prof_filter_n241.log 2>&1; echo "rc=$?"
date
ls -la $out | grep ${tag}_
