#!/bin/bash
# One gpurun call: parity tests, the bench lines and the ncu evidence kept under profiles/.
#   usage (on the GPU box): bash tools/measure_round.sh r01
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/${R}_pytest_gpu.txt 2>&1
tail -3 $O/${R}_pytest_gpu.txt
python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench_n1.err
# other BASELINE configs (parity-checked in the same run; not the headline)
python bench.py --frames 1024 --width 1920 --height 1080 --steps 5 --warmup 3 --no-cpu-baseline > $O/${R}_cfg2_1080p.json 2>> $O/${R}_bench_n1.err
python bench.py --frames 512 --gridn 14 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${R}_cfg4_n14_L0.json 2>> $O/${R}_bench_n1.err
python bench.py --frames 512 --gridn 14 --level 1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${R}_cfg4_n14_L1.json 2>> $O/${R}_bench_n1.err
python bench.py --frames 512 --gridn 14 --level 3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${R}_cfg4_n14_L3.json 2>> $O/${R}_bench_n1.err
# ncu: launch list of the bench command (shares, not absolutes), then full captures of K1 and K2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches.csv \
    python bench.py --frames 1024 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/${R}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chess_cascade -s 2 -c 1 -f -o $O/${R}_k1 \
    python bench.py --frames 512 --chunk 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/${R}_k1_ncu.log 2>&1
python tools/bench_blobs.py --frames 256 --steps 2 > $O/${R}_blobs_4k_n14.json 2>> $O/${R}_bench_n1.err
python tools/bench_blobs.py --frames 256 --steps 2 --kind circles --gridn 10 > $O/${R}_blobs_4k_circles.json 2>> $O/${R}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_blob_launches.csv \
    python tools/bench_blobs.py --frames 64 --chunk 64 --steps 1 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cluster_find -s 2 -c 1 -f -o $O/${R}_k2 \
    python bench.py --frames 512 --chunk 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/${R}_k2_ncu.log 2>&1
python tools/bench_levels.py > $O/${R}_levels_4k_n14.jsonl 2>> $O/${R}_bench_n1.err
python tools/sweep_resolutions.py > $O/${R}_cfg5_sweep.jsonl 2>> $O/${R}_bench_n1.err
python tools/bench_dense.py > $O/${R}_dense_4k.txt 2>> $O/${R}_bench_n1.err
python tools/bench_preproc.py > $O/${R}_preproc_4k.txt 2>> $O/${R}_bench_n1.err
python tools/bench_boards.py > $O/${R}_boards_4k.jsonl 2>> $O/${R}_bench_n1.err
python tools/bench_boards.py --gridn 10 --level 0 >> $O/${R}_boards_4k.jsonl 2>> $O/${R}_bench_n1.err
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${R}_gpu.txt
ls -la $O | tail -20
