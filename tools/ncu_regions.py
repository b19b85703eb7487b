"""Summarise an .ncu-rep: headline metrics plus executed instructions / stall samples per SASS region.
usage: python tools/ncu_regions.py gpurun_out/x.ncu-rep [chunk]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'launch__block_size', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__shared_mem_per_block_dynamic', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h and float(vals[i] or 0) > 0.05):
        print(f"{h:90s} {vals[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc = hdr.index('Source'); ie = hdr.index('Instructions Executed'); it = hdr.index('Thread Instructions Executed'); iss = hdr.index('# Samples')
data = [(r[isrc], int(r[ie] or 0), int(r[it] or 0), int(r[iss] or 0)) for r in rows[2:] if len(r) > it]
tot = sum(d[1] for d in data); tots = sum(d[3] for d in data)
print("total warp-inst", tot, "samples", tots, "sass lines", len(data))
for a in range(0, len(data), chunk):
    ch = data[a:a + chunk]
    e = sum(d[1] for d in ch); s = sum(d[3] for d in ch); t = sum(d[2] for d in ch)
    if e == 0 and s == 0: continue
    nv = sum(1 for d in ch if 'VABSDIFF4' in d[0]); nl = sum(1 for d in ch if 'LDG' in d[0]); ns = sum(1 for d in ch if 'SYNCS' in d[0])
    print(f"{a:5d} inst {e/tot*100:5.1f}% samples {s/tots*100:5.1f}% thr/inst {t/max(e,1):5.1f} vabs {nv} ldg {nl} syncs {ns}")
# dynamic opcode histogram: executed warp-instructions per SASS opcode (which pipe the issue slots go to)
import re, collections
hist = collections.Counter(); samp = collections.Counter()
for srcline, e, t, s in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", srcline)
    op = m.group(2) if m else "?"
    hist[op] += e; samp[op] += s
print("opcode            warp-inst     share   stall-samples share")
for op, e in hist.most_common(28):
    print(f"{op:14s} {e:12d} {e/max(tot,1)*100:8.2f}% {samp[op]/max(tots,1)*100:10.2f}%")
