import sys; sys.path.insert(0, '.')
import numpy as np
from mrgingham_b200 import api, synth
img = synth.board_frame(320, 240, 10, seed=3)
print(api.find_chessboard_corners_int(img, 0).shape)
