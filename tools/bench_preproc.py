"""Times the preprocessing kernels (blur, normalize+CLAHE) on device-resident 4K frames with CUDA events.
   python tools/bench_preproc.py [--frames 256]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from mrgingham_b200 import api, synth
    api._require_gpu()
    base = np.stack([synth.board_frame(a.width, a.height, 10, seed=s) for s in range(4)])
    frames = torch.from_numpy(base).cuda().repeat((a.frames + 3) // 4, 1, 1)[:a.frames].contiguous()
    det = api.Detector(max_frames=a.frames)
    out = torch.empty_like(frames)
    npx = a.frames * a.width * a.height
    for name, kw in (("blur r=1", dict(clahe=False, blur_radius=1)), ("blur r=2", dict(clahe=False, blur_radius=2)),
                     ("blur r=4", dict(clahe=False, blur_radius=4)), ("normalize+clahe", dict(clahe=True, blur_radius=0)),
                     ("normalize+clahe+blur r=1", dict(clahe=True, blur_radius=1))):
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            det.preprocess(frames, out=out, stream=s, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            det.preprocess(frames, out=out, stream=s, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print("%-26s %8.3f ms per %d %dx%d frames, %7.1f Gpix/s" % (name, ms, a.frames, a.width, a.height, npx / ms / 1e6))


if __name__ == "__main__":
    main()
