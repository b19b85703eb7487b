"""One summary file per kernel from an `ncu --set full` report that holds many kernels (tools/measure_round.sh):
profiles/<prefix>_<kernel>_ncu_summary.txt with the launch geometry, DRAM bytes, achieved GB/s against the measured
HBM peak, issue/pipe activity, residency and the leading stall reasons; plus <prefix>_kernels_ncu_table.txt.
usage: python tools/ncu_all_summary.py gpurun_out/x.ncu-rep profiles/r02"""
import csv, io, json, os, re, subprocess, sys
rep, prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
peak = 6538.3
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return 0.0


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return num(v) * m.get(unit, 1)


def to_us(v, unit):
    m = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
    return num(v) * m.get(unit, 1)


seen = {}
table = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]]
    short = re.sub(r"^void ", "", name).replace("mrgb200::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    short = re.sub(r"\((int|bool)\)", "", short)
    short = re.sub(r"\(.*", "", short)
    key = re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_")
    dur = to_us(r[col['gpu__time_duration.sum']], units[col['gpu__time_duration.sum']])
    # keep the longest instance of each kernel (the full-size launch)
    if key in seen and seen[key][0] >= dur:
        continue
    rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
    wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
    gbs = (rd + wr) / (dur * 1e-6) / 1e9 if dur > 0 else 0
    lines = [f"# ncu --set full --clock-control none, one launch of {short} (tools/exercise_all.py: 64 frames of 3840x2160 unless the kernel ran on a subset)",
             f"kernel                                   {name[:160]}",
             f"dram_bytes_total                         {rd + wr:.0f} byte",
             f"dram_achieved                            {gbs:.1f} GB/s = {gbs / peak:.3f} of the measured HBM peak ({peak} GB/s)"]
    for h in want:
        if h in col:
            lines.append(f"{h:90s} {r[col[h]]} {units[col[h]]}")
    stalls = [(num(r[i]), h) for h, i in col.items() if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
    for v, h in sorted(stalls, reverse=True)[:6]:
        lines.append(f"{h:90s} {v:.3f} inst")
    seen[key] = (dur, lines, short, gbs, r)
for key, (dur, lines, short, gbs, r) in sorted(seen.items()):
    open(f"{prefix}_{key}_ncu_summary.txt", "w").write("\n".join(lines) + "\n")
    table.append(f"{short:48s} {dur:10.1f} us  dram {gbs:7.1f} GB/s ({gbs / peak:5.3f})  issue {num(r[col['smsp__issue_active.avg.pct_of_peak_sustained_active']]):5.1f} %  "
                 f"warps {num(r[col['sm__warps_active.avg.pct_of_peak_sustained_active']]):5.1f} %  regs {r[col['launch__registers_per_thread']]}")
open(f"{prefix}_kernels_ncu_table.txt", "w").write("# every kernel of tools/exercise_all.py, one ncu --set full capture each (longest launch per kernel)\n" + "\n".join(table) + "\n")
print("\n".join(table))
