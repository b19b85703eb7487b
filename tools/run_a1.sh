#!/bin/bash
# first GPU call of the round: parity tests, then the new bench line
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/a1_pytest.txt 2>&1; tail -5 $O/a1_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/a1_bench.json 2> $O/a1_bench.err; tail -c 3000 $O/a1_bench.json; tail -5 $O/a1_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-overlap --no-e2e --no-cpu-baseline --no-content --sustain-seconds 0 > $O/a1_bench_noov.json 2>> $O/a1_bench.err; tail -c 600 $O/a1_bench_noov.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/a1_bench_ref.json 2>> $O/a1_bench.err; tail -c 600 $O/a1_bench_ref.json
