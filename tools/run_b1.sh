#!/bin/bash
# blob path after the per-border rewrite: parity tests, bench, chunk-size A/B, launch list, ncu of the walk kernel
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_blobs.py -x -q > $O/b1_pytest.txt 2>&1; tail -5 $O/b1_pytest.txt
for c in 32 64 128; do
  echo "chunk $c: $(MRG_B200_BLOB_CHUNK=$c timeout 300 python tools/bench_blobs.py --frames 1024 --steps 2 2>>$O/b1_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['frames_per_s'], d['ms_per_step'], d['kernel_ms_per_step'], d['parity'])")" | tee -a $O/b1_chunks.txt
done
timeout 300 python tools/bench_blobs.py --frames 1024 --steps 2 > $O/b1_blobs_4k_n14.json 2>> $O/b1_err.txt; cat $O/b1_blobs_4k_n14.json
timeout 300 python tools/bench_blobs.py --frames 1024 --steps 2 --kind circles --gridn 10 > $O/b1_blobs_4k_circles.json 2>> $O/b1_err.txt; cat $O/b1_blobs_4k_circles.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/b1_blob_launches.csv \
    python tools/bench_blobs.py --frames 64 --chunk 64 --steps 1 --warmup 1 > /dev/null 2>&1
grep -o '"unnamed>::[a-z_]*\|"gpu__time_duration.sum","ns","[0-9]*"' $O/b1_blob_launches.csv | paste - - | tail -5
tail -3 $O/b1_err.txt
