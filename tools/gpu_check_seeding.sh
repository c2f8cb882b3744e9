#!/bin/bash
# GPU box: the GPU parity tests, then an ncu capture of the seeding sweeps on the default workload (3.1 Gbp)
python -m pytest tests -m gpu -q > gpurun_out/r2_t23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t23.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r2_p23_plain.json 2> gpurun_out/r2_p23_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 4 -c 4 -o gpurun_out/r02_sweeps_3g_tab python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r2_p23_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
