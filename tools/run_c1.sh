#!/bin/bash
# whole GPU test suite, the mixed-resolution batch, and one ncu --set full capture of every kernel
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/c1_pytest.txt 2>&1; tail -6 $O/c1_pytest.txt
timeout 600 python tools/bench_mixed.py > $O/c1_mixed.json 2> $O/c1_err.txt; cat $O/c1_mixed.json; tail -3 $O/c1_err.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'pyramid|chess_|cluster_|blur|clahe|minmax|norm_lut|blob_' -c 80 -f -o $O/c1_all \
    python tools/exercise_all.py > $O/c1_all_ncu.log 2>&1
tail -2 $O/c1_all_ncu.log; ls -la $O/c1_all.ncu-rep
timeout 300 python tools/bench_latency.py > $O/c1_latency.jsonl 2>> $O/c1_err.txt; cat $O/c1_latency.jsonl
