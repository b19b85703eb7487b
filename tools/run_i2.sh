#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mixed_batch.py -m gpu -x -q > $O/i2_pytest.txt 2>&1; tail -3 $O/i2_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-content --sustain-seconds 0 --no-e2e > $O/i2_bench.json 2> $O/i2_err.txt
python -c "
import json
d=json.load(open('$O/i2_bench.json')); r=d['roofline']; print(round(d['value']/1e3), 'Gpix/s frac', round(r['frac'],4), 'k1 ms', round(r['avg_launch_ms'],3), 'k2 ms', round(r['k2_avg_launch_ms'],3), d['parity']['identical_to_oracle'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cluster_find -s 2 -c 1 -f -o $O/i2_k2 \
    python bench.py --frames 512 --chunk 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-content --sustain-seconds 0 --base-frames 8 --no-overlap > $O/i2_ncu.log 2>&1
timeout 300 python tools/bench_boards.py > $O/i2_boards.jsonl 2>> $O/i2_err.txt; cut -c1-200 $O/i2_boards.jsonl
timeout 300 python tools/bench_mixed.py > $O/i2_mixed.json 2>> $O/i2_err.txt; cut -c1-170 $O/i2_mixed.json
