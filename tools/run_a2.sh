#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/a2_pytest.txt 2>&1; tail -3 $O/a2_pytest.txt
bash tools/ab_k1.sh "" "MRG_B200_K1_SUMS=1" "MRG_B200_K1_SUMS=2" "MRG_B200_K1_MINB=6" \
  "MRG_B200_K1_NOCARRY=1 MRG_B200_K1_STAGES=3 MRG_B200_K1_MINB=6" "MRG_B200_K1_NOCARRY=1 MRG_B200_K1_STAGES=3 MRG_B200_K1_MINB=7" \
  "MRG_B200_K1_NOCARRY=1 MRG_B200_K1_STAGES=3 MRG_B200_K1_MINB=6 MRG_B200_K1_SUMS=1" "MRG_B200_K1_NOCARRY=1 MRG_B200_K1_STAGES=3 MRG_B200_K1_MINB=7 MRG_B200_K1_SUMS=1" \
  "MRG_B200_K1_NOCARRY=1 MRG_B200_K1_STAGES=3 MRG_B200_K1_MINB=7 MRG_B200_K1_SUMS=2" ""
cp $O/ab_k1.txt $O/a2_ab_k1.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chess_cascade -s 2 -c 1 -f -o $O/a2_k1 \
    python bench.py --frames 512 --chunk 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-content --sustain-seconds 0 --base-frames 8 --no-overlap > $O/a2_k1_ncu.log 2>&1
ls -la $O/a2_k1.ncu-rep
timeout 900 python bench.py --steps 20 --warmup 5 > $O/a2_bench.json 2> $O/a2_bench.err; tail -c 2500 $O/a2_bench.json; tail -5 $O/a2_bench.err
