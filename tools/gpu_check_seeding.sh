#!/bin/bash
# GPU box: seeding parity tests, the default bench (3.1 Gbp), and the same without the Bloom filters
python -m pytest tests -m gpu -q -x -k "seeding or digest or synthetic" > gpurun_out/r2_t29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t29.log
B200_DEBUG=1 timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/r2_b29.json 2> gpurun_out/r2_b29.log; echo "rc=$?" >> gpurun_out/r2_b29.log
B200_BLOOM=0 timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r2_b29_nobloom.json 2> gpurun_out/r2_b29_nobloom.log
