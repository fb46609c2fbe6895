"""Aggregate an ncu report's per-instruction counters by CUDA source line.

  python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [top] [samples]

Joins `ncu --page source --csv` (SASS rows with executed-instruction counts and
stall samples) with `nvdisasm -g` line markers of the in-tree libsphb200.so.
"""

import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jax_sph_b200", "libsphb200.so")


def sass_lines():
    """{mangled function: [(offset, file, line)]}"""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=tmp, check=True, capture_output=True,
                         text=True).stdout
    funcs, cur, loc = {}, None, ("?", 0)
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), loc[0], loc[1], m.group(2).strip()))
    return funcs


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    by_samples = len(sys.argv) > 4 and sys.argv[4] == "samples"  # order by stall samples
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and r:
            cur["rows"].append(r)
    funcs = sass_lines()
    srcs = {}
    for b in blocks:
        if want and want not in b["name"]:
            continue
        hdr, data = b["rows"][0], b["rows"][1:]
        ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index(
            "Thread Instructions Executed")
        # find the matching function by instruction count + first opcodes
        cand = None
        for name, ins in funcs.items():
            if len(ins) == len(data) and all(
                    ins[k][3].split()[0] == data[k][1].split()[0] for k in (0, 1, 2, len(data) // 2)):
                key = re.sub(r"[^A-Za-z0-9]", "", b["name"])
                if cand is None or sum(t in name for t in re.findall(r"[A-Z][a-z]+", b["name"])) > 0:
                    cand = name
        if cand is None:
            print("no SASS match for", b["name"][:100])
            continue
        ins = funcs[cand]
        per, tot, tots = {}, 0, 0
        for k, d in enumerate(data):
            n, s, t = int(d[ii]), int(d[si]), int(d[ti])
            key = (ins[k][1], ins[k][2])
            a = per.setdefault(key, [0, 0, 0])
            a[0] += n; a[1] += s; a[2] += t
            tot += n; tots += s
        print(f"== {b['name'][:110]}\n   matched {cand[:90]}\n   warp-inst {tot:,}  samples {tots:,}")
        for (fn, ln), (n, s, t) in sorted(per.items(), key=lambda kv: -kv[1][1 if by_samples else 0])[:top]:
            if fn not in srcs:
                p = os.path.join(ROOT, "jax_sph_b200", "csrc", fn)
                srcs[fn] = open(p).read().splitlines() if os.path.exists(p) else []
            text = srcs[fn][ln - 1].strip()[:95] if 0 < ln <= len(srcs[fn]) else ""
            print(f"   {n / tot * 100:5.1f}% inst {s / max(tots, 1) * 100:5.1f}% smp thr={t / max(n, 1):4.1f} "
                  f"{fn}:{ln:<4} {text}")


if __name__ == "__main__":
    main()
