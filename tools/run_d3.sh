#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_grid.py tests/test_grid_vs_ref.py -m gpu -x -q > $O/d3_pytest.txt 2>&1; tail -3 $O/d3_pytest.txt
timeout 600 python tools/bench_boards.py > $O/d3_boards.jsonl 2> $O/d3_err.txt; cut -c1-200 $O/d3_boards.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/d3_board_launches.csv \
    python tools/bench_boards.py --frames 64 --chunk 64 --steps 1 --warmup 1 > /dev/null 2>&1
grep cluster_refine $O/d3_board_launches.csv | grep -o '"ns","[0-9]*"' | tr '\n' ' '
