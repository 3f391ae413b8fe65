#!/bin/bash
# K2 iteration on the GPU box: parity tests of the element kernels, A/B timing of variant builds.
tag=${1:-k2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_tests.txt 2>&1
echo "gpu tests rc=$?" | tee -a gpurun_out/${tag}_tests.txt
tail -4 gpurun_out/${tag}_tests.txt
for lib in hybrid_fem_lssvr_b200/libhfl.so $EXTRA_LIBS; do
  HFL_LIB=$lib timeout 300 python scripts/probe_variant.py 2>&1 | tee -a gpurun_out/${tag}_probe.txt
done
