#!/bin/bash
# GPU box: seeding parity tests, the default bench (3.1 Gbp), forward sweeps compiled for 8 / 6 blocks per SM
python -m pytest tests -m gpu -q -x -k "seeding or digest or synthetic or odd" > gpurun_out/r2_t31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t31.log
B200_DEBUG=1 timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/r2_b31.json 2> gpurun_out/r2_b31.log; echo "rc=$?" >> gpurun_out/r2_b31.log
for m in 8 6; do B200_FWD_MINB=$m timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r2_b31_f$m.json 2> gpurun_out/r2_b31_f$m.log; done
