"""Hottest SASS instructions of the (one) kernel in an .ncu-rep, by stall samples, with the stall reasons of each and the
CUDA source line it came from. usage: python tools/ncu_hot_lines.py x.ncu-rep [top]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 70
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
print("# columns:", [h for h in hdr][:60])
isrc = hdr.index('Source'); ie = hdr.index('Instructions Executed'); iss = hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and not h.endswith('_not_issued')]
data = []
for k, r in enumerate(rows[2:]):
    if len(r) <= iss: continue
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols if i < len(r)), reverse=True)[:3]
    data.append((k, r[isrc], int(r[ie] or 0), int(r[iss] or 0), st))
tots = sum(d[3] for d in data); tote = sum(d[2] for d in data)
print(f"# total samples {tots}, warp instructions {tote}")
tot_reason = collections.Counter()
for k, r in enumerate(rows[2:]):
    for i in stall_cols:
        if i < len(r): tot_reason[hdr[i][6:]] += int(r[i] or 0)
print("# stall reasons over the kernel:", ", ".join(f"{n} {v/max(tots,1)*100:.1f}%" for n, v in tot_reason.most_common(12)))
for k, s, e, n, st in sorted(data, key=lambda d: -d[3])[:top]:
    print(f"{k:5d} samples {n/tots*100:5.2f}% inst {e/tote*100:5.2f}%  {' '.join(f'{nm}:{v}' for v, nm in st if v):48s} {s.strip()[:90]}")
# context of the 6 hottest: the 6 instructions before each
print("# context")
hot = sorted(data, key=lambda d: -d[3])[:8]
for k, s, e, n, st in hot:
    print(f"--- around {k}")
    for d in data[max(0, k - 8):k + 3]:
        print(f"   {d[0]:5d} smp {d[3]:6d} ex {d[2]:9d}  {d[1].strip()[:100]}")
# shared-memory wavefronts: ideal against actual, per load/store instruction (bank conflicts)
try:
    iw = hdr.index('L1 Wavefronts Shared'); ii = hdr.index('L1 Wavefronts Shared Ideal')
    tw = ti = 0; per = []
    for k, r in enumerate(rows[2:]):
        if len(r) <= max(iw, ii): continue
        w = float(r[iw] or 0); i_ = float(r[ii] or 0)
        if w > 0: per.append((w, i_, k, int(r[ie] or 0), r[isrc].strip()[:70])); tw += w; ti += i_
    print(f"# shared-memory wavefronts: {tw:.0f} against {ti:.0f} ideal ({tw / max(ti, 1):.2f}x)")
    for w, i_, k, e, s in sorted(per, reverse=True)[:24]:
        print(f"{k:5d} wavefronts {w:12.0f} ideal {i_:12.0f} ({w / max(i_, 1):.2f}x) executed {e:9d}  {s}")
except ValueError:
    pass
