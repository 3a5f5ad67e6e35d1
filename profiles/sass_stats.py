"""SASS of the hot kernels, as shipped: per kernel the instruction listing (encodings stripped) and an opcode histogram,
whole kernel and innermost hot loop (the backward branch with the largest body that contains the kernel's gathers).

    python profiles/sass_stats.py [out_dir=profiles/r02/sass]

Run here (no GPU needed): cuobjdump reads the sm_100a cubins inside mdsea_b200/_lib/libmdsea_b200.so.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mdsea_b200", "_lib", "libmdsea_b200.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02", "sass")
KERNELS = {   # label -> substring of the mangled name
    "k_tile_force_f32_LJ": "k_tile_force_f32ILi1E",
    "k_tile_build": "k_tile_build",
    "k_nbr_force_f32_LJ_tex_cs": "k_nbr_force_f32ILb1ELi1ELb1E",
    "k_nbr_build_warp_expanded_2groups": "k_nbr_build_warpILb1ELi2E",
    "k_allpairs_n3_f64_3d_LJ_pbc": "k_allpairs_n3_f64ILi3ELi1ELb0ELb1E",
    "k_allpairs_n3_f32_3d_LJ_pbc": "k_allpairs_n3_f32ILi3ELi1ELb0ELb1E",
    "k_cut_kick_finish_fused": "k_cut_kick_finishILb1E",
}

dump = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = {}, None
for ln in dump.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", ln):
        funcs[cur].append(ln)

os.makedirs(OUT, exist_ok=True)
summary = []
for label, key in KERNELS.items():
    name = next((f for f in funcs if key in f), None)
    if name is None:
        summary.append(f"{label}: not found ({key})")
        continue
    ins = []
    for ln in funcs[name]:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))

    def opcode(t):
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        return t.split()[0].split(".")[0]

    hist = collections.Counter(opcode(t) for _, t in ins)
    # innermost hot loop: backward branches; pick the one whose body holds the most gathers (LDS.128 / LDG / TLD / broadcast LDS)
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            body = [x for x in ins if lo <= x[0] <= a]
            score = sum(1 for _, x in body if x.startswith(("LDS", "LDG", "TLD")) or " LDS" in x or " LDG" in x) * 1000 - len(body)
            if len(body) > 40 and (best is None or score > best[0]):
                best = (score, lo, a, body)
    with open(os.path.join(OUT, label + ".sass"), "w") as f:
        f.write(f"// {name}\n// cuobjdump -sass of mdsea_b200/_lib/libmdsea_b200.so (sm_100a), encodings stripped\n")
        for a, t in ins:
            f.write(f"/*{a:04x}*/ {t}\n")
    line = f"{label}: {len(ins)} instructions; " + ", ".join(f"{k} {v}" for k, v in hist.most_common(14))
    if best:
        _, lo, hi, body = best
        h2 = collections.Counter(opcode(t) for _, t in body)
        line += f"\n    hot loop /*{lo:04x}*/../*{hi:04x}*/: {len(body)} instructions; " + ", ".join(f"{k} {v}" for k, v in h2.most_common(16))
    summary.append(line)
with open(os.path.join(OUT, "SUMMARY.txt"), "w") as f:
    f.write("# opcode histograms of the hot kernels (profiles/sass_stats.py); full listings beside this file\n")
    f.write("\n".join(summary) + "\n")
print("\n".join(summary))
