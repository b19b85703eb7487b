"""CPU-only, world_size 2 over gloo: the multi-process plumbing of the N>1 path -- contiguous
sharding of the batch with no data-path collective, the max-over-ranks reduction bench.py uses for
timing, and a frame-order gather of per-rank results."""
import os
import socket

import numpy as np
import pytest

from mrgingham_b200.sharding import owner_of, shard_range


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 8, 100, 4096):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                covered += list(range(lo, hi))
            assert covered == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert owner_of(0, 4096, 8) == 0 and owner_of(4095, 4096, 8) == 7 and owner_of(512, 4096, 8) == 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nframes, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(nframes, rank, world)
    # stand-in for the per-frame detector output of this rank's shard: (frame index, a count)
    local = torch.tensor([[i, 100 + (i % 3)] for i in range(lo, hi)], dtype=torch.int64).reshape(-1, 2)
    # no collective on the data path; results are gathered (tiny) in frame order afterwards
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo], dtype=torch.int64))
    gathered = [torch.zeros((int(s.item()), 2), dtype=torch.int64) for s in sizes]
    dist.all_gather(gathered, local) if len({int(s.item()) for s in sizes}) == 1 else None
    # the timing reduction bench.py does: max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((float(t.item()), [int(s.item()) for s in sizes],
               torch.cat(gathered).numpy() if len({int(s.item()) for s in sizes}) == 1 else None))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_reduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    nframes, world = 64, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, sizes, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 11.0
    assert sizes == [32, 32]
    assert np.array_equal(gathered[:, 0], np.arange(nframes))


def test_bench_reference_arm_contract():
    """bench.py --impl reference: the reference's own CPU code on the host cores, one JSON line with the keys the
    driver reads (same metric/unit/config as the GPU arm, impl, cpu_baseline, e2e with zero copies)"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--width", "640", "--height", "480", "--base-frames", "2"],
                         check=True, capture_output=True, text=True, timeout=120).stdout.strip().split("\n")
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpix/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] and "workload" in line["config"]
