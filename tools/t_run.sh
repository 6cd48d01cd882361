#!/bin/bash
# Dev harness (GPU box): accuracy of the folded kernels per configuration, timings of the
# transposed strip kernels against the one-tile-per-TMEM kernels, then the GPU test suite.
mkdir -p gpurun_out
timeout 300 python tools/fold_check.py fwd > gpurun_out/t_fwd.log 2>&1; echo "fwd rc $?"
timeout 300 python tools/fold_check.py inv > gpurun_out/t_inv.log 2>&1; echo "inv rc $?"
timeout 300 python tools/fold_check.py bench > gpurun_out/t_bench.log 2>&1; echo "bench rc $?"
grep -c "ok " gpurun_out/t_fwd.log gpurun_out/t_inv.log; grep -h "BAD\|Error\|error" gpurun_out/t_fwd.log gpurun_out/t_inv.log | head -20
cat gpurun_out/t_bench.log | tail -12
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_t.log 2>&1; echo "pytest rc $?"
tail -5 gpurun_out/pytest_t.log
BRV_TC_VARIANT=4 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v4.log 2>&1; echo "pytest (variant 4) rc $?"
tail -3 gpurun_out/pytest_v4.log
