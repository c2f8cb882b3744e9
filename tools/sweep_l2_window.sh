#!/bin/bash
# tuning sweep (GPU box): L2 access-policy window over the occ table (B200_L2_WINDOW) at three index sizes
for m in 1 0; do B200_L2_WINDOW=$m python bench.py --ref-bp 100000000 --steps 12 --warmup 5 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r2_b20_100m_l2w$m.json 2> gpurun_out/r2_b20_100m_l2w$m.log; done
B200_L2_WINDOW=0 python bench.py --steps 12 --warmup 5 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r2_b20_3g_l2w0.json 2> gpurun_out/r2_b20_3g_l2w0.log
