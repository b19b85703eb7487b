#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: Mpix/s of ChESS + clustering ("NMS") over
synthetic 4K chessboard frames (BASELINE.json metric), on N GPUs of one node.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code

A "step" is one pass of the detector over one batch: BASELINE.json configs[2], 4096 frames of 3840x2160, 10x10
board, level 0, PER GPU (weak scaling: images are independent, every rank runs the same-sized batch on its own
frames, no collective on the data path; `--scaling strong` shards ONE 4096-frame batch over the ranks instead).
Consecutive passes are software-pipelined over two detectors: pass p+1's ChESS kernels start while pass p's last
clustering launch and result copies finish (`--no-overlap`: collect each pass before the next is enqueued).
Prints ONE JSON line (rank 0); exits non-zero WITHOUT a value if any rank's corner lists differ from the oracle's.
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames in one batch (per GPU with --scaling weak, in all with strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--gridn", type=int, default=10)
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--base-frames", type=int, default=64, help="distinct synthetic frames, tiled to --frames")
    ap.add_argument("--chunk", type=int, default=2048, help="frames per kernel launch (at least two launches per rank are made)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="collect every pass before enqueueing the next")
    ap.add_argument("--no-content", action="store_true", help="skip roofline.by_content (K1 on other frame contents)")
    ap.add_argument("--sustain-seconds", type=float, default=3.0, help="length of the extra sustained-clock run (0 = skip)")
    ap.add_argument("--kernel-variant", type=int, default=0)
    ap.add_argument("--max-points", type=int, default=256, help="per-frame output capacity (the boards have gridn^2 corners)")
    return ap.parse_args()


def workload_name(a):
    return f"{a.frames} x {a.width}x{a.height} grayscale, {a.gridn}x{a.gridn} board, level {a.level}"


def bench_config(a, world):
    """identical for both arms (the driver compares them): what is processed, not how"""
    per_gpu = a.frames if a.scaling == "weak" else None
    return {"workload": workload_name(a) + (" per GPU" if a.scaling == "weak" and world > 1 else ""),
            "frames_total": a.frames * world if a.scaling == "weak" else a.frames,
            "frames_per_gpu": per_gpu if per_gpu is not None else f"{a.frames}/{world}",
            "distinct_frames": max(1, min(a.base_frames, a.frames)),
            "l2": "inputs (>= 4 GB per GPU) far larger than the 126 MB L2; no flush needed",
            "parallelism": f"independent frames over {world} GPU(s), {a.scaling} scaling, no collective on the data path"}


def _gen_frame(args):
    from mrgingham_b200 import synth
    w, h, n, seed = args
    return synth.board_frame(w, h, n, seed=seed)


def make_base_frames(a, world=1):
    """K distinct boards (seeds 0..K-1), rendered on a few host processes (0.7 s each at 4K on one core).
    Must run before CUDA is initialised in this process (fork)."""
    K = max(1, min(a.base_frames, a.frames))
    jobs = [(a.width, a.height, a.gridn, s) for s in range(K)]
    procs = max(1, min(K, (os.cpu_count() or 1) // max(world, 1), 16))
    if procs == 1 or K <= 2:
        return [_gen_frame(j) for j in jobs]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_gen_frame, jobs)


CONTENT_KINDS = (
    ("clean", "clean board (sigma 2 noise, 3x3 blur: the bench frames)", None),
    ("noisy", "board + sigma 6 noise, no blur", None),
    ("textured", "board on a blurred-noise background (sigma 2 noise, 3x3 blur)", None),
    ("clahe", "board through --clahe then --blur 1 on the GPU (detector config clahe=1, blur_radius=1)", dict(clahe=True, blur_radius=1)),
    ("checker", "dense checker, period 8", None),
)
CONTENT_DISTINCT = 4


def _gen_content(args):
    from mrgingham_b200 import synth
    kind, w, h, n, s = args
    if kind == "clean":
        return synth.board_frame(w, h, n, seed=100 + s)
    if kind == "noisy":
        return synth.board_frame(w, h, n, seed=100 + s, noise_sigma=6.0, blur=False)
    if kind == "clahe":
        return synth.board_frame(w, h, n, seed=100 + s, blur=False)
    if kind == "checker":
        return synth.checker_frame(w, h, 8, seed=300 + s)
    # textured: the board composited over low-contrast blurred noise, then the usual sensor noise and blur
    b0 = synth.board_frame(w, h, n, seed=100 + s, noise_sigma=0.0, blur=False)
    bg = (synth.blurred_noise_frame(w, h, seed=200 + s, passes=3) // 2 + 64).astype(np.uint8)
    f = np.where(b0 == synth.BACKGROUND, bg, b0).astype(np.float32)
    f += np.random.default_rng(400 + s).normal(0.0, 2.0, size=f.shape).astype(np.float32)
    return synth.box_blur3(np.clip(np.rint(f), 0, 255).astype(np.uint8))


def make_content_frames(a):
    """frames for roofline.by_content, rendered on a few host processes before CUDA is initialised"""
    jobs = [(k, a.width, a.height, a.gridn, s) for k, _, _ in CONTENT_KINDS for s in range(CONTENT_DISTINCT)]
    procs = max(1, min(len(jobs), os.cpu_count() or 1, 16))
    if procs == 1:
        out = [_gen_content(j) for j in jobs]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            out = pool.map(_gen_content, jobs)
    return {k: out[i * CONTENT_DISTINCT:(i + 1) * CONTENT_DISTINCT] for i, (k, _, _) in enumerate(CONTENT_KINDS)}


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

    def nvml_window(self, t0, t1):
        """clocks between two time.time() stamps from the NVML samples, or None if there are none"""
        inside = [r for r in list(self.nvml_rows) if t0 <= r[0] <= t1]
        if not inside:
            return None
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        reasons = sorted({name for r in inside for name, bit in bits if r[2] & bit})
        mhz = [r[1] for r in inside]
        return {"sm_mhz": statistics.median(mhz), "sm_mhz_min": min(mhz), "sm_max_mhz": self.nvml_max, "samples": len(inside),
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
            if self.t0 is not None and self.t1 is not None:
                nvml = self.nvml_window(self.t0, self.t1)
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
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": max(r[2] for r in rows) if rows else None,
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    base = make_base_frames(a)          # the same frames as our arm
    # each step = a bounded sample of the workload: ~2 s of all-core CPU work over the base frames, cycled
    for _ in range(a.warmup):
        cpu_throughput(base, a.level, 0.5, cores)
    vals, nframes, secs, kind = [], 0, 0.0, "port"
    for _ in range(a.steps):
        v, kind, n, dt = cpu_throughput(base, a.level, 2.0, cores)
        vals.append(v); nframes += n; secs += dt
    value = nframes * a.width * a.height / secs / 1e6
    sample = (f"{nframes} frames of {a.width}x{a.height} (the bench's {len(base)} base frames, cycled) in {secs:.1f} s over "
              f"{a.steps} steps, {cores} threads, whole frames per thread")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * secs / max(a.steps, 1), "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": bench_config(a, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(gpu_index):
    """Pin this process to the CPU cores NVML reports as local to the GPU, BEFORE any pinned host memory is allocated
    (first touch then puts the staging pools on the GPU's NUMA node). Returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return "no NVML affinity inside this process's cpuset"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} cores local to GPU {gpu_index} (of {len(allowed)} allowed)"
    except Exception as e:      # no NVML, no permission: run unbound
        return f"unbound ({type(e).__name__})"


def run_ours(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    base = make_base_frames(a, world)       # before CUDA is up in this process (the pool forks)
    K = len(base)
    content = make_content_frames(a) if world == 1 and not a.no_content else None

    import torch
    import torch.distributed as dist
    from mrgingham_b200 import api, synth

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    affinity = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    W, H = a.width, a.height
    from mrgingham_b200.sharding import shard_range
    if a.scaling == "weak":
        lo, hi = rank * a.frames, (rank + 1) * a.frames        # every rank its own full-size batch
        total_frames = a.frames * world
    else:
        lo, hi = shard_range(a.frames, rank, world)            # contiguous shard: frame i -> rank floor(i*world/frames)
        total_frames = a.frames
    nloc = hi - lo

    # synthetic data: K distinct frames, tiled; frame i of the job is base[i % K]
    base_t = torch.from_numpy(np.stack(base)).to(dev)
    frames = torch.empty((nloc, H, W), dtype=torch.uint8, device=dev)
    for i in range(nloc):
        frames[i].copy_(base_t[(lo + i) % K])
    del base_t
    torch.cuda.synchronize()

    # at least two launches per rank, so that the clustering kernel of one chunk overlaps the ChESS kernel of the next
    chunk = max(1, min(a.chunk, max(64, (nloc + 1) // 2)))
    ndet = 1 if a.no_overlap else 2
    dets = [api.Detector(max_frames=chunk, max_rows=H, max_cols=W, max_points=a.max_points, device=local_rank,
                         kernel_variant=a.kernel_variant) for _ in range(ndet)]
    det = dets[0]
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    class Counters:
        k1_ms = k1_n = k2_ms = k2_n = other_n = 0

    def run_passes(npasses, cnt=None):
        """npasses passes over the batch; pass p runs on detector p % ndet and is collected after pass p+1 has been
        enqueued (ndet == 2), so its clustering tail and result copies overlap the next pass's ChESS kernels.
        Returns the results of the last pass."""
        out = None
        inflight = []
        for p in range(npasses):
            d = dets[p % ndet]
            d.enqueue(frames, a.level, stream=stream)
            inflight.append(d)
            if len(inflight) == ndet:
                d0 = inflight.pop(0)
                out = d0.collect()
                if cnt is not None:
                    ms, n = d0.last_kernel_ms(0); cnt.k1_ms += ms; cnt.k1_n += n
                    ms, n = d0.last_kernel_ms(1); cnt.k2_ms += ms; cnt.k2_n += n
                    ms, n = d0.last_kernel_ms(2); cnt.other_n += n
        for d0 in inflight:
            out = d0.collect()
            if cnt is not None:
                ms, n = d0.last_kernel_ms(0); cnt.k1_ms += ms; cnt.k1_n += n
                ms, n = d0.last_kernel_ms(1); cnt.k2_ms += ms; cnt.k2_n += n
                ms, n = d0.last_kernel_ms(2); cnt.other_n += n
        return out

    sampler = ClockSampler(local_rank)
    sampler.start()
    if a.warmup > 0:
        run_passes(a.warmup)

    # parity gate in the same run, on EVERY rank: every frame's corner list must equal the oracle's for its base frame.
    # One untimed pass per detector is made for it whatever --warmup says.
    from oracle import pyoracle as po
    want = [po.find_corners(b, a.level) for b in base]
    ok = True
    for d in dets:
        xy, counts = d.find_corners(frames, a.level, stream=stream)
        for i in range(nloc):
            wnt = want[(lo + i) % K]
            if counts[i] != len(wnt) or not np.array_equal(xy[i, :counts[i]], wnt):
                ok = False
                break
    ok_all = allreduce(1.0 if ok else 0.0, dist.ReduceOp.MIN if world > 1 else None) > 0.5
    parity = {"frames_checked": nloc * ndet * world, "distinct_frames": K, "ranks_checked": world, "identical_to_oracle": bool(ok_all),
              "corners_per_frame": int(len(want[0]))}
    if not ok_all:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "error": "PARITY FAILED: corner lists "
                              "differ from the oracle's on at least one rank; no throughput is reported", "parity": parity}), flush=True)
        sampler.stop()
        if world > 1:
            dist.destroy_process_group()
        sys.exit(3)

    # ---- timed region: device-resident inputs, exactly --steps passes, CUDA events on the launching stream
    for d in dets:
        d.set_profiling(True)
    cnt = Counters()
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_passes(a.steps, cnt)      # the last collect() has waited for the last clustering launch and result copy ...
    e1.record()                   # ... so this event closes over all of the device work of the K steps
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if a.sustain_seconds <= 0 else None
    elapsed_ms = allreduce(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None)
    ms_per_step = elapsed_ms / a.steps
    value = total_frames * W * H / (ms_per_step * 1e-3) / 1e6

    # ---- the same loop for >= --sustain-seconds: does the number hold at the clock the part sustains?
    sustained = None
    if a.sustain_seconds > 0:
        npass = max(a.steps, int(np.ceil(a.sustain_seconds * 1e3 / ms_per_step)))
        scnt = Counters()
        barrier()
        t0 = time.time()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        run_passes(npass, scnt)
        s1.record()
        barrier()
        t1 = time.time()
        sms = allreduce(s0.elapsed_time(s1), dist.ReduceOp.MAX if world > 1 else None)
        win = sampler.nvml_window(t0, t1) or {}
        sustained = {"passes": npass, "seconds": sms * 1e-3, "value": total_frames * W * H / (sms / npass * 1e-3) / 1e6, "unit": UNIT,
                     "k1_avg_launch_ms": scnt.k1_ms / max(scnt.k1_n, 1),
                     "sm_mhz_median": win.get("sm_mhz"), "sm_mhz_min": win.get("sm_mhz_min"), "clock_samples": win.get("samples"),
                     "reasons": win.get("reasons")}
        clocks = sampler.stop()
    for d in dets:
        d.set_profiling(False)

    # ---- e2e: host (pinned) frames through the C ABI, H2D + D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        pool_n = min(nloc, chunk, 512)     # 4.2 GB of pinned host memory per pool
        npools = 2 if ndet == 2 else 1     # two pools: the copies of call c+1 start while call c's tail is collected
        pools = []
        for k in range(npools):
            pool = torch.empty((pool_n, H, W), dtype=torch.uint8).pin_memory()
            for i in range(pool_n):
                pool[i].copy_(torch.from_numpy(base[(lo + k * pool_n + i) % K]))
            pools.append(pool)
        calls = (nloc + pool_n - 1) // pool_n

        def e2e_step():
            left = nloc
            inflight = []
            for c in range(calls):
                n = min(pool_n, left)
                d = dets[c % ndet]
                d.enqueue(pools[c % npools][:n], a.level, stream=stream)
                inflight.append(d)
                if len(inflight) == ndet:
                    inflight.pop(0).collect()
                left -= n
            for d in inflight:
                d.collect()

        e2e_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(a.e2e_steps):
            e2e_step()
        s1.record()
        barrier()
        ems = allreduce(s0.elapsed_time(s1), dist.ReduceOp.MAX if world > 1 else None)
        # what the box can do: the same pinned pools, copied host->device and nothing else, all ranks at once
        scratch = torch.empty((pool_n, H, W), dtype=torch.uint8, device=dev)
        scratch.copy_(pools[0], non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for c in range(calls):
            scratch.copy_(pools[c % npools], non_blocking=True)
        c1.record()
        barrier()
        cms = allreduce(c0.elapsed_time(c1), dist.ReduceOp.MAX if world > 1 else None)
        del scratch
        e2e = {"value": total_frames * W * H / (ems / a.e2e_steps * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(total_frames) * W * H,
               "d2h_bytes_per_step": int(total_frames) * (a.max_points * 2 * 4 + 8),
               "h2d_ceiling": {"value": world * calls * pool_n * W * H / (cms * 1e-3) / 1e6, "unit": UNIT,
                               "what": "pinned host->device copies of the same pools and nothing else, all ranks at once"},
               "cpu_affinity": affinity,
               "note": f"host-pinned frames via mrg_b200_find_corners_batch_enqueue/collect, {pool_n} frames per call, {npools} pinned pool(s) per rank"}
        del pools

    # ---- K1 on other frame contents (one GPU only): the cascade's throughput depends on edge / noise density
    by_content = None
    if world == 1 and not a.no_content:
        by_content = content_table(a, content, po, torch, dev)

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
    frames_per_launch = nloc * a.steps / max(cnt.k1_n, 1)
    lvl_px = (W >> a.level) * (H >> a.level) if a.level else W * H
    bytes_per_launch = frames_per_launch * lvl_px          # algorithmic: 1 byte per pixel read
    avg_launch_ms = cnt.k1_ms / max(cnt.k1_n, 1)
    achieved = bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
    traffic = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_pixel"] * bytes_per_launch
            traffic_src = f"{tj['dram_bytes_per_pixel']:.3f} B/px from {tj.get('source', 'profiles/k1_traffic.json')}, scaled to this launch's pixels"
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "chess_cascade_kernel (K1)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "avg_launch_ms": avg_launch_ms, "bytes_per_launch": bytes_per_launch, "frames_per_launch": frames_per_launch,
                "k1_share_of_step": cnt.k1_ms / elapsed_ms,
                # the same bytes over the whole timed region: the fraction of the HBM peak the WHOLE detector runs at. When
                # k1_share_of_step is well above 1 (short launches: the K1 launches of the two detectors run concurrently),
                # the per-launch `frac` above understates K1 and this is the number to read.
                "step_frac": (nloc * a.steps * lvl_px / (elapsed_ms * 1e-3) / 1e9) / peak,
                "k2_avg_launch_ms": cnt.k2_ms / max(cnt.k1_n, 1),
                "sustained_frac": (frames_per_launch * lvl_px / (sustained["k1_avg_launch_ms"] * 1e-3) / 1e9 / peak) if sustained and sustained["k1_avg_launch_ms"] > 0 else None,
                "by_content": by_content}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, kind, n, dt = cpu_throughput(base, a.level, 12.0, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n} frames of {W}x{H} (the bench's {K} base frames, cycled) in {dt:.1f} s, {cores} threads, whole frames per thread"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": bench_config(a, world),
        "launch": {"frames_per_launch": chunk, "detectors": ndet,
                   "passes": "overlapped: pass p is collected after pass p+1 is enqueued" if ndet == 2 else "collected one by one"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(cnt.k1_n + cnt.k2_n + cnt.other_n),
        "clocks": clocks, "sustained": sustained, "parity": parity,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def content_table(a, content, po, torch, dev):
    """K1's roofline fraction on 4K frames of different content (64 frames = 4 distinct x 16, two launches each; the
    corner lists of the distinct frames are checked against the oracle). The cascade only ever does MORE work than on
    the clean board: every flagged 8-pixel cell costs ~8x an unflagged one."""
    from mrgingham_b200 import api
    W, H = a.width, a.height
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    nd, reps = CONTENT_DISTINCT, 16
    rows = []
    for key, name, cfg in CONTENT_KINDS:
        fr = content[key]
        d = api.Detector(max_frames=32, max_rows=H, max_cols=W, max_points=4096, candidate_capacity=1 << 21, device=dev.index,
                         kernel_variant=a.kernel_variant, **(cfg or {}))
        t = torch.from_numpy(np.stack([fr[i % nd] for i in range(nd * reps)])).to(dev)
        d.set_profiling(True)
        # K1 alone (mrg_b200_chess_candidates_batch): its candidate counts against the oracle's dense response
        ms = []
        for _ in range(4):
            ncand, _ = d.chess_candidates(t, a.level)
            ms.append(d.last_kernel_ms(0)[0])
        k1 = float(np.median(ms[1:]))
        ok = None
        if not cfg and a.level == 0:     # the preprocessing chain has its own parity tests (tests/test_preproc.py, test_blur.py)
            want = [int((po.chess_response_5(f, fill=0)[7:-7, 7:-7] > 15).sum()) for f in fr]
            ok = all(int(ncand[i]) == want[i % nd] for i in range(nd * reps))
        # the whole detector only where the candidate lists are of a size the clustering kernel is built for
        corners = None
        if int(ncand.max()) <= 20000:
            xy, counts = d.find_corners(t, a.level)
            corners = int(counts[0])
            if not cfg:
                wantc = [po.find_corners(f, a.level) for f in fr]
                ok = bool(ok) and all(counts[i] == len(wantc[i % nd]) and np.array_equal(xy[i, :counts[i]], wantc[i % nd])
                                      for i in range(nd * reps))
        gbs = nd * reps * W * H / (k1 * 1e-3) / 1e9
        rows.append({"content": name, "k1_ms": k1, "achieved": gbs, "frac": gbs / peak, "corners": corners,
                     "candidates_per_frame": int(np.median(ncand)), "identical_to_oracle": bool(ok) if ok is not None else None})
        d.close()
        del t
        torch.cuda.empty_cache()
    return rows


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
