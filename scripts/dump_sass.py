"""SASS listings of the hot kernels of the in-tree libsphb200.so -> profiles/<prefix>_sass_<name>.txt
(cuobjdump -sass, one function per file, with a mnemonic histogram at the top).

  python scripts/dump_sass.py [prefix]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jax_sph_b200", "libsphb200.so")
WANT = {
    "k_duo_search_3d": r"k_duoILi3ENS_8PhysNoneELi1ELi1EEE",
    "k_duo_density_filter_3d_qsk": r"k_duoILi3ENS_11PhysDensityILi3ELi0ELi0EEELi3ELi1EEE",
    "k_duo_force_3d_qsk_tvf_uniform_eta": r"k_duoILi3ENS_9PhysForceILi3ELi0ELi0ELi3EEELi2ELi2EEE",
    "k_duo_force_3d_qsk_tvf": r"k_duoILi3ENS_9PhysForceILi3ELi0ELi0ELi1EEELi2ELi2EEE",
}


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        if cur is not None:
            cur.append(ln)
    for name, pat in WANT.items():
        hit = [k for k in funcs if re.search(pat, k)]
        if not hit:
            print("no function matches", pat)
            continue
        body = funcs[hit[0]]
        ops = collections.Counter()
        for ln in body:
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                ops[m.group(1).split(".")[0] if not m.group(1).startswith(("UBLKCP", "SYNCS")) else m.group(1)] += 1
        path = os.path.join(ROOT, "profiles", f"{prefix}_sass_{name}.txt")
        with open(path, "w") as f:
            f.write(f"# cuobjdump -sass jax_sph_b200/libsphb200.so, function {hit[0]}\n")
            f.write(f"# {sum(ops.values())} instructions; mnemonics: " +
                    ", ".join(f"{k} {v}" for k, v in ops.most_common()) + "\n")
            # (instruction encodings dropped: the listing is for reading)
            lines = [re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", ln).rstrip() for ln in body]
            f.write("\n".join(ln for ln in lines if ln.strip()) + "\n")
        print(path, sum(ops.values()), "instructions", {k: v for k, v in ops.items() if k.startswith(("UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2"))})


if __name__ == "__main__":
    main()
