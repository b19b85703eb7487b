#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_grid.py tests/test_grid_vs_ref.py tests/test_gpu_api_contract.py tests/test_gpu_dropin_pymodule.py tests/test_gpu_blobs.py -m gpu -x -q > $O/d4_pytest.txt 2>&1; tail -3 $O/d4_pytest.txt
timeout 600 python tools/bench_boards.py > $O/d4_boards.jsonl 2> $O/d4_err.txt; timeout 300 python tools/bench_boards.py --gridn 10 --level 0 >> $O/d4_boards.jsonl 2>> $O/d4_err.txt
timeout 300 python tools/bench_boards.py --frames 1024 >> $O/d4_boards.jsonl 2>> $O/d4_err.txt
timeout 300 python tools/bench_boards.py --host >> $O/d4_boards.jsonl 2>> $O/d4_err.txt
cut -c1-330 $O/d4_boards.jsonl; tail -3 $O/d4_err.txt
