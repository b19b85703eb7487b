#!/bin/bash
# One gpurun call: smoke, parity tests, the bench lines and the ncu evidence kept under profiles/.
#   usage (on the GPU box): bash tools/measure_round.sh r02      (then copy gpurun_out/<R>_* and gpurun_out/ncu/* to profiles/)
R=${1:-r02}
O=gpurun_out
mkdir -p $O $O/ncu
python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.txt 2>&1; tail -1 $O/${R}_smoke.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/${R}_pytest_gpu.txt 2>&1
tail -3 $O/${R}_pytest_gpu.txt
timeout 900 python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench_n1.err
# other BASELINE configs (parity-checked in the same run; not the headline)
timeout 600 python bench.py --frames 1024 --width 1920 --height 1080 --steps 5 --warmup 3 --no-cpu-baseline --no-content --sustain-seconds 0 > $O/${R}_cfg2_1080p.json 2>> $O/${R}_bench_n1.err
timeout 600 python bench.py --frames 512 --gridn 14 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-content --sustain-seconds 0 > $O/${R}_cfg4_n14_L0.json 2>> $O/${R}_bench_n1.err
timeout 600 python tools/bench_mixed.py > $O/${R}_cfg5_mixed_batch.json 2>> $O/${R}_bench_n1.err
# ncu: launch list of the bench command (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches.csv \
    python bench.py --frames 1024 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-content --sustain-seconds 0 > $O/${R}_launches_bench.log 2>&1
# one ncu --set full capture of every kernel, summarised here (the report itself is too large to bring back)
timeout 1500 ncu --set full --clock-control none -k regex:'pyramid|chess_|cluster_|blur|clahe|minmax|norm_lut|blob_|gather|16' -c 120 -f -o $O/${R}_all \
    python tools/exercise_all.py > $O/${R}_all_ncu.log 2>&1
python tools/ncu_all_summary.py $O/${R}_all.ncu-rep $O/ncu/${R} > $O/${R}_kernels_table.txt 2>&1; rm -f $O/${R}_all.ncu-rep
# K1 alone, with source-level counters: summarised here (regions, opcode mix, hottest SASS lines with their stall reasons)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chess_cascade -s 2 -c 1 -f -o $O/${R}_k1 \
    python bench.py --frames 512 --chunk 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-content --sustain-seconds 0 --base-frames 8 --no-overlap > $O/${R}_k1_ncu.log 2>&1
python tools/ncu_regions.py $O/${R}_k1.ncu-rep > $O/${R}_k1_ncu_summary.txt 2>&1
python tools/ncu_hot_lines.py $O/${R}_k1.ncu-rep 90 > $O/${R}_k1_hot_lines.txt 2>&1; rm -f $O/${R}_k1.ncu-rep
timeout 300 python tools/bench_blobs.py --frames 1024 --steps 2 > $O/${R}_blobs_4k_n14.json 2>> $O/${R}_bench_n1.err
timeout 300 python tools/bench_blobs.py --frames 1024 --steps 2 --kind circles --gridn 10 > $O/${R}_blobs_4k_circles.json 2>> $O/${R}_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_blob_launches.csv \
    python tools/bench_blobs.py --frames 64 --chunk 64 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 300 python tools/bench_levels.py > $O/${R}_levels_4k_n14.jsonl 2>> $O/${R}_bench_n1.err
timeout 300 python tools/bench_dense.py > $O/${R}_dense_4k.txt 2>> $O/${R}_bench_n1.err
timeout 300 python tools/bench_preproc.py > $O/${R}_preproc_4k.txt 2>> $O/${R}_bench_n1.err
timeout 300 python tools/bench_boards.py > $O/${R}_boards_4k.jsonl 2>> $O/${R}_bench_n1.err
timeout 300 python tools/bench_boards.py --frames 1024 >> $O/${R}_boards_4k.jsonl 2>> $O/${R}_bench_n1.err
timeout 300 python tools/bench_boards.py --gridn 10 --level 0 >> $O/${R}_boards_4k.jsonl 2>> $O/${R}_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_board_launches.csv \
    python tools/bench_boards.py --frames 64 --chunk 64 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 300 python tools/bench_latency.py > $O/${R}_latency.jsonl 2>> $O/${R}_bench_n1.err
timeout 900 bash tools/run_sanitizer.sh > $O/${R}_sanitizer.txt 2>&1
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${R}_gpu.txt
tail -5 $O/${R}_bench_n1.err; cat $O/${R}_kernels_table.txt; ls $O | wc -l
