#!/bin/bash
# K1 iteration on the GPU box: parity tests of the coarse solve, timing probe, ncu of the three kernels.
# usage (under gpurun): bash scripts/gpu_k1.sh [tag]
tag=${1:-k1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fem.py tests/test_gpu_general.py -x -q -m gpu > gpurun_out/${tag}_tests.txt 2>&1
echo "fem tests rc=$?" | tee -a gpurun_out/${tag}_tests.txt
tail -5 gpurun_out/${tag}_tests.txt
timeout 300 python scripts/probe_k1.py hybrid_fem_lssvr_b200/libhfl.so $EXTRA_LIBS > gpurun_out/${tag}_probe.txt 2>&1
cat gpurun_out/${tag}_probe.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fem_ -c 5 -o gpurun_out/${tag}_ncu -f \
    env N_ROUNDS=1 python scripts/probe_k1.py hybrid_fem_lssvr_b200/libhfl.so > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
