#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cluster_find -s 2 -c 1 -f -o $O/i1_k2 \
    python bench.py --frames 512 --chunk 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-content --sustain-seconds 0 --base-frames 8 --no-overlap > $O/i1_ncu.log 2>&1
ls -la $O/i1_k2.ncu-rep
