#!/bin/bash
# GPU box: seeding-related parity tests, then the default bench (3.1 Gbp) with the backward sweeps compiled for 6 and 9 blocks per SM
python -m pytest tests -m gpu -q -x -k "seeding or single_job or digest or synthetic or odd or alt" > gpurun_out/r2_t24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t24.log
B200_DEBUG=1 timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/r2_b24.json 2> gpurun_out/r2_b24.log; echo "rc=$?" >> gpurun_out/r2_b24.log
B200_BWD_MINB=9 timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r2_b24_b9.json 2> gpurun_out/r2_b24_b9.log; echo "rc=$?" >> gpurun_out/r2_b24_b9.log
