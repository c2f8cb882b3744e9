#!/bin/bash
# GPU box: launch list of a short default run, then ncu --set full of the four seeding sweeps and of the DP / CIGAR kernels
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r02f_plain.json 2> gpurun_out/r02f_plain.log || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r02f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 4 -c 4 -o gpurun_out/r02f_sweeps_3g python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r02f_ncu_sweeps.log 2>&1
ls -la gpurun_out/*.ncu-rep
