"""Per-source-line share of executed instructions and stall samples of ONE kernel in an ncu report.
ncu's source page is SASS-level in CSV form; the line of every SASS address comes from `nvdisasm -g` of the cubin
inside the object file the library was linked from (compiled with -lineinfo).
usage: python profiles/hot_lines.py report.ncu-rep kernel_substring [object.o] [top_n]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kname = sys.argv[1], sys.argv[2]
obj = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mdsea_b200/_lib/obj/cells.o")
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the report may hold several kernels: blocks start with a "Kernel Name" row
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], rows=[])
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
blk = next(b for b in blocks if kname in b["name"])
hdr = blk["rows"][0]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
isrc = hdr.index("Source")
base = None
counts = {}
for r in blk["rows"][1:]:
    if len(r) <= isamp or not r[ia]:
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    counts[a - base] = (int(r[ii] or 0), int(r[isamp] or 0), r[isrc])
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
mangled = None
line_of, cur_line, cur_file, on = {}, None, None, False
for ln in dis:
    if ln.startswith(".text."):
        on = kname in ln
        continue
    if ln.startswith("\t.section") or ln.startswith("//-----"):
        if ".text." in ln and kname not in ln:
            on = False
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur_file, cur_line)
by_line = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for off, (n, s, _) in counts.items():
    k = line_of.get(off, ("?", 0))
    by_line[k][0] += n
    by_line[k][1] += s
    tot_i += n
    tot_s += s
srcs = {}
print(f"# {blk['name'][:90]}\n# {tot_i} warp instructions, {tot_s} stall samples")
print(f"{'instr%':>7s} {'samples%':>8s}  file:line")
for (f, l), (n, s) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    p = os.path.join(os.path.dirname(os.path.abspath(obj)), "../../csrc", f or "")
    if f and os.path.exists(p):
        if f not in srcs:
            srcs[f] = open(p).read().splitlines()
        if 0 < l <= len(srcs[f]):
            text = srcs[f][l - 1].strip()[:100]
    print(f"{100 * n / max(tot_i, 1):7.2f} {100 * s / max(tot_s, 1):8.2f}  {f}:{l}  {text}")
