#!/bin/bash
# tuning sweep (GPU box): share of the SM the persistent seeding sweeps take vs end-to-end throughput with chunks sharing the device
run() { tag=$1; shift; env "$@" python bench.py --ref-bp 1000000000 --steps 8 --warmup 5 --no-cpu-baseline --parity-chunks 0 > gpurun_out/r2_b16_$tag.json 2> gpurun_out/r2_b16_$tag.log; }
run fill100 B200_SEED_FILL=1.0
run fill78 B200_SEED_FILL=0.78
run fill67 B200_SEED_FILL=0.67
run fill50 B200_SEED_FILL=0.5
