#!/bin/bash
# 16-bit preprocessing + parallel refinement: whole GPU suite, board bench, refine kernel time
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/d1_pytest.txt 2>&1; tail -6 $O/d1_pytest.txt
timeout 600 python tools/bench_boards.py > $O/d1_boards.jsonl 2> $O/d1_err.txt; timeout 300 python tools/bench_boards.py --gridn 10 --level 0 >> $O/d1_boards.jsonl 2>> $O/d1_err.txt
cat $O/d1_boards.jsonl; tail -3 $O/d1_err.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/d1_board_launches.csv \
    python tools/bench_boards.py --frames 64 --chunk 64 --steps 1 --warmup 1 > /dev/null 2>&1
grep -o '"mrgb200::[a-z_0-9]*\|unnamed>::[a-z_0-9]*\|"gpu__time_duration.sum","ns","[0-9]*"' $O/d1_board_launches.csv | paste - - | awk '{n[$1]++; gsub(/[^0-9]/,"",$2); t[$1]+=$2} END {for (k in n) printf "%-40s %4d launches %10.3f ms\n", k, n[k], t[k]/1e6}' | sort -k4 -n -r | head -12
