#!/bin/bash
# final_run.sh -- the round-end sequence on one GPU box (gpurun -- 'bash tools/final_run.sh'): GPU tests, smoke, the default bench line,
# the reference arm, the ncu launch list of the bench command and one --set full capture of the interpreter kernel; everything lands in
# gpurun_out/final_*, from where the summaries under profiles/ are made (tools/ncu_summary.py, tools/ncu_sass_regions.py)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:icw:: -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 3 -c 1 -o gpurun_out/final_c3 -f python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu --no-workloads --no-parity > gpurun_out/final_c3_ncu.log 2>&1
cat gpurun_out/final_tests.log; tail -3 gpurun_out/final_smoke.log
