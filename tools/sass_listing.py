"""SASS evidence per production kernel: profiles/sass/<kernel>.sass (full listing of the production instantiation,
gzip'd when long) and profiles/sass/opcodes.txt (static opcode histogram of every kernel: UTMALDG = TMA,
VABSDIFF4 / IDP.4A = the byte-lane arithmetic, SYNCS = mbarriers). usage: python tools/sass_listing.py"""
import collections, gzip, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mrgingham_b200", "libmrgingham_b200.so")
out = os.path.join(ROOT, "profiles", "sass"); os.makedirs(out, exist_ok=True)
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
archs = set(re.findall(r"arch = (sm_\w+)", txt))
lines = [f"# cuobjdump -sass mrgingham_b200/libmrgingham_b200.so: {len(parts) - 1} kernels, cubin architectures {sorted(archs)}",
         "# static instruction counts per kernel (instantiation); full listings beside this file"]
keep_full = ("chess_cascade_kernel<3, true>", "chess_tiled_kernel<true, true>", "cluster_find_kernel", "blob_walk_kernel", "box_blur3_kernel", "pyramid_l1_kernel")
for p in parts[1:]:
    mangled = p.split("\n", 1)[0].strip()
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"^void ", "", name).replace("(anonymous namespace)::", "").replace("mrgb200::", "")
    short = re.sub(r"\(.*", "", short)
    ops = collections.Counter(m.group(1).split(".")[0] for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", p))
    total = sum(ops.values())
    lead = ", ".join(f"{k} {v}" for k, v in ops.most_common(10))
    flags = " ".join(f"{k}={ops[k]}" for k in ("UTMALDG", "SYNCS", "VABSDIFF4", "IDP", "LDGSTS", "ATOMS", "SHFL", "VOTE") if ops.get(k))
    lines.append(f"{short:44s} {total:6d} instr | {flags} | {lead}")
    if any(k in short for k in keep_full):
        fn = re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_") + ".sass.gz"
        with gzip.open(os.path.join(out, fn), "wt") as f:
            f.write("Function : " + p)
open(os.path.join(out, "opcodes.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
