#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_dp_lanes.py -m gpu -q 2>&1 | grep -v "^$" > $O/r02_t12.log; grep -n "^E  .*Error\|^E   \|^FAILED\|passed\|failed" $O/r02_t12.log | head -40
