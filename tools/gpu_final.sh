#!/bin/bash
# GPU box: the whole GPU test suite, the default bench line (3.1 Gbp, with the CPU baseline), the configs[1] line (100 Mbp), the
# launch list of a short default run and an ncu --set full capture of the four seeding sweeps.  Outputs under gpurun_out/r02f_*.
python -m pytest tests -m gpu -q > gpurun_out/r02f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_tests.log
B200_DEBUG=1 timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1_3100mbp.json 2> gpurun_out/r02f_bench_n1_3100mbp.log; echo "rc=$?" >> gpurun_out/r02f_bench_n1_3100mbp.log
timeout 900 python bench.py --ref-bp 100000000 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_n1_100mbp.json 2> gpurun_out/r02f_bench_n1_100mbp.log; echo "rc=$?" >> gpurun_out/r02f_bench_n1_100mbp.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r02f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 4 -c 4 -o gpurun_out/r02f_sweeps_3g python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r02f_ncu_sweeps.log 2>&1
ls -la gpurun_out/*.ncu-rep
