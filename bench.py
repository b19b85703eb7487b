#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: Mpix/s of ChESS + clustering ("NMS") over
synthetic 4K chessboard frames (BASELINE.json metric), on N GPUs of one node.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code

A "step" is one pass of the detector over the whole batch (default: BASELINE.json configs[2],
4096 frames of 3840x2160, 10x10 board, level 0; strong scaling: the batch is sharded over the
ranks with no collective on the data path). Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpix/s ChESS+NMS over 4K frames at 1/2/4/8 GPU; achieved HBM GB/s vs peak"
UNIT = "Mpix/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="total frames in the batch (all ranks)")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--gridn", type=int, default=10)
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--base-frames", type=int, default=8, help="distinct synthetic frames, tiled to --frames")
    ap.add_argument("--chunk", type=int, default=2048, help="frames per kernel launch (at least two launches per rank are made)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-variant", type=int, default=0)
    ap.add_argument("--max-points", type=int, default=256, help="per-frame output capacity (the boards have gridn^2 corners)")
    return ap.parse_args()


def workload_name(a):
    return f"{a.frames} x {a.width}x{a.height} grayscale, {a.gridn}x{a.gridn} board, level {a.level}"


# ---------------------------------------------------------------------------------------------
# clocks sampling (the recipe's nvidia-smi line), during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load. Started before the
    warm-up so that the sampler is already running when the (short) timed region begins; samples
    are attributed to the timed region by timestamp, falling back to every sample taken under load
    (utilization >= 50 %) if the region was shorter than the sampling period."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.t0 = self.t1 = None
        # second source, for timed regions shorter than nvidia-smi's sampling period (small shards at 8 GPUs):
        # NVML polled every 4 ms from a thread of this process
        self.nvml_rows = []
        self.nvml_max = None
        self._nvml_stop = threading.Event()
        self._nvml_thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        try:
            self._nvml_thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self._nvml_thread.start()
        except Exception:
            self._nvml_thread = None

    def _nvml_loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._nvml_stop.is_set():
                t = time.time()
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                except Exception:
                    mask = 0
                self.nvml_rows.append((t, mhz, mask))
                time.sleep(0.004)
        except Exception:
            pass

    def _nvml_result(self):
        """clocks over the timed region from the NVML samples, or None if there are none inside it"""
        if self.t0 is None or self.t1 is None:
            return None
        inside = [r for r in list(self.nvml_rows) if self.t0 <= r[0] <= self.t1]
        if not inside:
            return None
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        reasons = sorted({name for r in inside for name, bit in bits if r[2] & bit})
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": self.nvml_max, "samples": len(inside),
                "scope": "timed region (NVML, 4 ms period)", "reasons": reasons}

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        self._nvml_stop.set()
        nvml = None
        try:
            nvml = self._nvml_result()
        except Exception:
            nvml = None
        if self.proc is None:
            return nvml or {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(p[1]), float(p[2]), float(p[4]), [v.lower().startswith("active") for v in p[5:9]]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= self.t1 + 0.05]
        scope = "timed region"
        if len(inside) < 3 and nvml is not None:
            return nvml          # the region was shorter than nvidia-smi's period: the NVML samples taken inside it
        if len(inside) < 3:
            inside = [r for r in rows if r[3] >= 50.0]
            scope = "warm-up + timed region (samples under load)"
        reasons = set()
        for r in inside:
            for name, act in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if act:
                    reasons.add(name)
        sm = [r[1] for r in inside]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(r[2] for r in rows) if rows else None,
                "samples": len(sm), "scope": scope, "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref) or, where it was not built, the oracle port
# ---------------------------------------------------------------------------------------------
def cpu_find_fn():
    from oracle import pyoracle as po
    if po.have_ref():
        po.ref_lib()
        return po.ref_find_corners, "reference"
    po.oracle_lib()
    return po.find_corners, "port"


def cpu_throughput(frames, level, min_seconds, threads):
    """Mpix/s of the CPU implementation with `threads` host threads, each looping over whole
    frames (the reference CLI's -j model, mrgingham-from-image.cc:50,374-379)."""
    fn, kind = cpu_find_fn()
    h, w = frames[0].shape
    done = [0] * threads
    stop = threading.Event()

    def worker(t):
        i = t
        while not stop.is_set():
            fn(frames[i % len(frames)], level)
            done[t] += 1
            i += threads

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    time.sleep(min_seconds)
    stop.set()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    n = sum(done)
    return n * w * h / dt / 1e6, kind, n, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mrgingham_b200 import synth
    cores = os.cpu_count() or 1
    base = [synth.board_frame(a.width, a.height, a.gridn, seed=s) for s in range(min(a.base_frames, 4))]
    # each step = a bounded sample of the workload: ~2 s of all-core CPU work
    for _ in range(a.warmup):
        cpu_throughput(base, a.level, 0.5, cores)
    vals, nframes, secs, kind = [], 0, 0.0, "port"
    for _ in range(a.steps):
        v, kind, n, dt = cpu_throughput(base, a.level, 2.0, cores)
        vals.append(v); nframes += n; secs += dt
    value = nframes * a.width * a.height / secs / 1e6
    sample = f"{nframes} frames of {a.width}x{a.height} in {secs:.1f} s over {a.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * secs / max(a.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "CPU arm: bounded sample of the same frames, all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from mrgingham_b200 import api, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    W, H = a.width, a.height
    # contiguous shard of the batch for this rank: frame i -> rank floor(i*world/frames)
    from mrgingham_b200.sharding import shard_range
    lo, hi = shard_range(a.frames, rank, world)
    nloc = hi - lo

    # synthetic data: K distinct frames, tiled; frame i of the batch is base[i % K]
    K = max(1, min(a.base_frames, a.frames))
    base = [synth.board_frame(W, H, a.gridn, seed=s) for s in range(K)]
    base_t = torch.from_numpy(np.stack(base)).to(dev)
    frames = torch.empty((nloc, H, W), dtype=torch.uint8, device=dev)
    for i in range(nloc):
        frames[i].copy_(base_t[(lo + i) % K])
    torch.cuda.synchronize()

    # at least two launches per rank, so that the clustering kernel of one chunk overlaps the ChESS kernel of the next
    chunk = max(1, min(a.chunk, max(64, (nloc + 1) // 2)))
    det = api.Detector(max_frames=chunk, max_rows=H, max_cols=W, max_points=a.max_points, device=local_rank,
                       kernel_variant=a.kernel_variant)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        det.enqueue(frames, a.level, stream=stream)
        return det.collect()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(a.warmup, 0)):
        xy, counts = step()

    # parity gate in the same run: every frame's corner list must equal the oracle's for its base frame
    parity = None
    if rank == 0:
        from oracle import pyoracle as po
        want = [po.find_corners(b, a.level) for b in base]
        ok = True
        for i in range(nloc):
            wnt = want[(lo + i) % K]
            if counts[i] != len(wnt) or not np.array_equal(xy[i, :counts[i]], wnt):
                ok = False
                break
        parity = {"frames_checked": nloc, "identical_to_oracle": ok, "corners_per_frame": int(len(want[0]))}

    # ---- timed region: device-resident inputs
    det.set_profiling(True)
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_ms = k1_n = k2_ms = k2_n = 0
    e0.record()
    for _ in range(a.steps):
        step()
        ms, n = det.last_kernel_ms(0); k1_ms += ms; k1_n += n
        ms, n = det.last_kernel_ms(1); k2_ms += ms; k2_n += n
        ms, n = det.last_kernel_ms(2); k2_n += n
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    det.set_profiling(False)
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / a.steps
    value = a.frames * W * H / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host (pinned) frames through the C ABI, H2D + D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        pool_n = min(nloc, chunk, 512)     # 4.2 GB of pinned host memory per call
        pool = torch.empty((pool_n, H, W), dtype=torch.uint8).pin_memory()
        for i in range(pool_n):
            pool[i].copy_(torch.from_numpy(base[(lo + i) % K]))
        calls = (nloc + pool_n - 1) // pool_n

        def e2e_step():
            left = nloc
            for _ in range(calls):
                n = min(pool_n, left)
                det.find_corners(pool[:n], a.level, stream=stream)
                left -= n

        e2e_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(a.e2e_steps):
            e2e_step()
        s1.record()
        barrier()
        ems = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([ems], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": a.frames * W * H / (ems / a.e2e_steps * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(a.frames) * W * H,
               "d2h_bytes_per_step": int(a.frames) * (a.max_points * 2 * 4 + 8),
               "note": f"host-pinned frames via mrg_b200_find_corners_batch, {pool_n} frames per call"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1: ChESS + candidate emission)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    frames_per_launch = nloc * a.steps / max(k1_n, 1)
    lvl_px = (W >> a.level) * (H >> a.level) if a.level else W * H
    bytes_per_launch = frames_per_launch * lvl_px          # algorithmic: 1 byte per pixel read
    avg_launch_ms = k1_ms / max(k1_n, 1)
    achieved = bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_pixel"] * bytes_per_launch
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "chess_sparse (K1)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": avg_launch_ms, "bytes_per_launch": bytes_per_launch,
                "k1_share_of_step": k1_ms / elapsed_ms if world == 1 else None,
                "k2_avg_launch_ms": k2_ms / max(k1_n, 1)}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, kind, n, dt = cpu_throughput(base, a.level, 12.0, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n} frames of {W}x{H} (the bench's base frames, cycled) in {dt:.1f} s, {cores} threads, whole frames per thread"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_name(a), "frames_per_gpu": nloc, "frames_per_launch": chunk,
                   "distinct_frames": K, "l2": "inputs (>= 4 GB per GPU) far larger than the 126 MB L2; no flush needed",
                   "parallelism": f"batch sharded over {world} GPU(s), no collective"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(k1_n + k2_n),
        "clocks": clocks, "parity": parity,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
