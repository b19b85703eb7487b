#!/bin/bash
O=gpurun_out; mkdir -p $O
cat > /tmp/t.py <<'PY'
import numpy as np
from mrgingham_b200 import api, synth
img = synth.board_frame(300, 200, 6, seed=1)
print(len(api.find_chessboard_corners_int(img, 0)))
img = synth.board_frame(1920, 1080, 10, seed=1)
print(len(api.find_chessboard_corners_int(img, 0)))
PY
PYTHONPATH=$PWD timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/t.py > $O/dbg_sanitizer.txt 2>&1
head -60 $O/dbg_sanitizer.txt
