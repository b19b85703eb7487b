#!/bin/bash
# one ncu --set full capture of every kernel, summarised ON the box (the report itself is too large to bring back)
O=gpurun_out; mkdir -p $O
timeout 1500 ncu --set full --clock-control none -k regex:'pyramid|chess_|cluster_|blur|clahe|minmax|norm_lut|blob_|gather' -c 80 -f -o $O/c2_all \
    python tools/exercise_all.py > $O/c2_all_ncu.log 2>&1
tail -2 $O/c2_all_ncu.log; ls -la $O/c2_all.ncu-rep
mkdir -p $O/ncu
python tools/ncu_all_summary.py $O/c2_all.ncu-rep $O/ncu/r02 > $O/c2_table.txt 2>&1; cat $O/c2_table.txt
rm -f $O/c2_all.ncu-rep
