import os, sys
sys.path.insert(0, os.getcwd())
import torch, numpy as np, ctypes
from mrgingham_b200 import api, synth
base = np.stack([synth.board_frame(3840,2160,10,seed=s) for s in range(2)])
n=64
x = torch.from_numpy(base).cuda().repeat(n//2,1,1)
out = torch.zeros((n,2160,3840), dtype=torch.int16, device='cuda')
det = api.Detector(max_frames=n)
L = api.lib()
def run():
    rc = L.mrg_b200_chess_response_batch(det._h, x.data_ptr(), 1, n, 2160, 3840, 3840, 3840*2160, out.data_ptr(), 1, None)
    assert rc == 0
run(); torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): run()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/3
px=n*3840*2160
print("dense response: %.3f ms per %d 4K frames = %.1f Gpix/s = %.0f GB/s (3 B/px)"%(ms,n,px/ms/1e6,3*px/ms/1e6))
from oracle import pyoracle as po
want = po.chess_response_5(base[1], fill=0)
print("parity", np.array_equal(out[1].cpu().numpy(), want))
