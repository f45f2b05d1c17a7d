#!/bin/bash
# kernel-only duration of the IS tensor-core kernel (one ncu pass, no replay of other metrics)
set -e
N=${1:-2000}
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "is_logpx" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:is_tc_kernel -s 1 -c 1 python tools/time_is.py $N 5000 bf16 2>&1 | grep -E "gpu__time|sm__cycles|sm__pipe|bf16:"
