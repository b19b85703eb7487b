#!/bin/bash
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/g1_bench_n2.json 2> $O/g1_err.txt; tail -c 1800 $O/g1_bench_n2.json; tail -3 $O/g1_err.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/g1_ref_n2.json 2>> $O/g1_err.txt; tail -c 600 $O/g1_ref_n2.json
