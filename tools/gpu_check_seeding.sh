#!/bin/bash
# GPU box: all GPU parity tests, the default bench (3.1 Gbp), and the launch list of a short run
python -m pytest tests -m gpu -q > gpurun_out/r2_t25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t25.log
B200_DEBUG=1 timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu-baseline > gpurun_out/r2_b25.json 2> gpurun_out/r2_b25.log; echo "rc=$?" >> gpurun_out/r2_b25.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches25.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r2_ncu25.log 2>&1
