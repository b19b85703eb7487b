#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_grid.py tests/test_grid_vs_ref.py tests/test_gpu_dropin_pymodule.py -m gpu -x -q > $O/d2_pytest.txt 2>&1; tail -4 $O/d2_pytest.txt
timeout 600 python tools/bench_boards.py > $O/d2_boards.jsonl 2> $O/d2_err.txt; timeout 300 python tools/bench_boards.py --gridn 10 --level 0 >> $O/d2_boards.jsonl 2>> $O/d2_err.txt
cut -c1-400 $O/d2_boards.jsonl; tail -3 $O/d2_err.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/d2_board_launches.csv \
    python tools/bench_boards.py --frames 64 --chunk 64 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/d2_bench.json 2>> $O/d2_err.txt; tail -c 1500 $O/d2_bench.json
