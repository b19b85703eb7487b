O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chess_cascade -s 2 -c 1 -f -o $O/k1hot python bench.py --frames 512 --chunk 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-content --sustain-seconds 0 --base-frames 8 --no-overlap > $O/k1hot.log 2>&1
python tools/ncu_hot_lines.py $O/k1hot.ncu-rep 90 > $O/k1_hot_lines.txt 2>&1
rm -f $O/k1hot.ncu-rep
head -5 $O/k1_hot_lines.txt | cut -c1-600
