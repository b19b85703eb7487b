#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_blobs.py tests/test_gpu_parity.py -m gpu -x -q > $O/f1_pytest.txt 2>&1; tail -3 $O/f1_pytest.txt
for c in 64 128; do
  echo "chunk $c: $(MRG_B200_BLOB_CHUNK=$c timeout 300 python tools/bench_blobs.py --frames 1024 --steps 2 2>>$O/f1_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['frames_per_s'], d['ms_per_step'], d['kernel_ms_per_step'], d['parity'])")" | tee -a $O/f1_chunks.txt
done
timeout 300 python tools/bench_blobs.py --frames 1024 --steps 2 > $O/f1_blobs_4k_n14.json 2>> $O/f1_err.txt; cut -c1-260 $O/f1_blobs_4k_n14.json
timeout 300 python tools/bench_blobs.py --frames 1024 --steps 2 --kind circles --gridn 10 > $O/f1_blobs_4k_circles.json 2>> $O/f1_err.txt; cut -c1-260 $O/f1_blobs_4k_circles.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/f1_blob_launches.csv \
    python tools/bench_blobs.py --frames 64 --chunk 64 --steps 1 --warmup 1 > /dev/null 2>&1
grep -o 'unnamed>::[a-z_]*\|"gpu__time_duration.sum","ns","[0-9]*"' $O/f1_blob_launches.csv | paste - - | tail -6
timeout 300 python tools/bench_latency.py > $O/f1_latency.jsonl 2>> $O/f1_err.txt; grep blobs $O/f1_latency.jsonl | cut -c1-220
tail -3 $O/f1_err.txt
