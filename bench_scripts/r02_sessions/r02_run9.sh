#!/bin/bash
cd "$(dirname "$0")/../.." && mkdir -p gpurun_out && O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" > $O/r02_t14.log; grep -n "^E  .*Error\|^E   .*assert\|^FAILED\|passed\|failed" $O/r02_t14.log | head -40
