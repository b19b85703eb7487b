#!/bin/bash
# blob path after the per-border rewrite: parity tests, bench, launch list
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_blobs.py -x -q > $O/b1_pytest.txt 2>&1; tail -15 $O/b1_pytest.txt
timeout 300 python tools/bench_blobs.py --frames 256 --steps 3 > $O/b1_blobs_4k_n14.json 2> $O/b1_err.txt; cat $O/b1_blobs_4k_n14.json; tail -3 $O/b1_err.txt
timeout 300 python tools/bench_blobs.py --frames 256 --steps 3 --kind circles --gridn 10 > $O/b1_blobs_4k_circles.json 2>> $O/b1_err.txt; cat $O/b1_blobs_4k_circles.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/b1_blob_launches.csv \
    python tools/bench_blobs.py --frames 64 --chunk 64 --steps 1 --warmup 1 > /dev/null 2>&1
grep -o '"unnamed>::[a-z_]*\|"gpu__time_duration.sum","ns","[0-9]*"' $O/b1_blob_launches.csv | paste - - | tail -12
