#!/bin/bash
# Round-end consolidation on the GPU box: full GPU test suite, smoke, bench line, ncu launch list of the bench step,
# ncu --set full of the hot kernels (scripts/profile_r02.py).  usage (under gpurun): bash scripts/gpu_final.sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/final_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'fem_|lssvr_element|dual_' -o gpurun_out/final_full -f \
    python scripts/profile_r02.py > gpurun_out/final_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
