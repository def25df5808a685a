#!/bin/bash
# One gpurun call: GPU parity tests, default bench, reference arm, launch list and ncu --set full captures.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_capture.sh <tag>'
# Outputs land in gpurun_out/<tag>_*; summaries are copied by hand into profiles/.
tag=${1:-r01b}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> $out/${tag}_gpu.txt

echo "== pytest -m gpu" ; date
timeout -k 5 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log

echo "== smoke" ; date
timeout -k 5 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log

echo "== bench (default)" ; date
timeout -k 5 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
head -c 1500 $out/${tag}_bench.json

echo "== bench (cold start, --prefill-rows 0)" ; date
timeout -k 5 300 python bench.py --prefill-rows 0 --no-cpu-baseline --kinship-rows 0 > $out/${tag}_bench_cold.json 2> $out/${tag}_bench_cold.err; echo "rc=$?"

echo "== bench --impl reference" ; date
timeout -k 5 600 python bench.py --impl reference --steps 3 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"
cat $out/${tag}_bench_ref.json | head -c 600

echo "== launch list" ; date
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-buffers 1 > $out/${tag}_launches_bench.log 2>&1; echo "rc=$?"

echo "== ncu full: filter, pair kernel, kinship" ; date
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:kg_scan_filter -s 24 -c 1 -f -o $out/${tag}_prof_filter \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --kinship-rows 0 --e2e-buffers 1 --cold-steps 0 > $out/${tag}_prof_filter.log 2>&1; echo "rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:kg_scan_pair -s 24 -c 1 -f -o $out/${tag}_prof_pair \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --kinship-rows 0 --e2e-buffers 1 > $out/${tag}_prof_pair.log 2>&1; echo "rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:kg_kinship_tc -s 2 -c 1 -f -o $out/${tag}_prof_kinship \
    python bench.py --steps 2 --warmup 3 --prefill-rows 0 --rows-per-step 1048576 --no-cpu-baseline > $out/${tag}_prof_kinship.log 2>&1; echo "rc=$?"
date
ls -la $out
