#!/bin/bash
# GPU box: extension-related parity tests, then the default bench with the DP lanes refilled at 8 (default), 1, 16 and 32 idle lanes
python -m pytest tests -m gpu -q -x -k "extend or synthetic or odd or wrappers or non_default" > gpurun_out/r2_t26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t26.log
timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/r2_b26_r8.json 2> gpurun_out/r2_b26_r8.log; echo "rc=$?" >> gpurun_out/r2_b26_r8.log
for r in 1 16 32; do B200_EXT_REFILL=$r timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r2_b26_r$r.json 2> gpurun_out/r2_b26_r$r.log; done
