"""Writes profiles/sass_r02.txt: per shipped kernel instantiation the instruction count, the counts of the mnemonics that
matter (TMA bulk copies UBLKCP, mbarrier SYNCS, MUFU seeds, 256-bit gathers LDG.E.ENL2.256, streaming list loads LDG.E.NA,
FP64 pipe instructions) and one example line of each -- from cuobjdump -sass of the objects linked into libnbody_b200.so.
  python profiles/sass_summary.py            (after the build; no GPU needed)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O = os.path.join(ROOT, "nbodysimulator.jl_b200", "csrc", "_build")
WANT = {"nbx_sympairs.o": ["sym_kernel"],
        "nbx_cells.o": ["verlet_force_kernel", "verlet_build_kernel", "cell_pairs2_kernel", "vv_pos_lists_kernel"],
        "nbx_allpairs.o": ["allpairs_kernel"],
        "nbx_slab.o": ["slab_ll_send_kernel", "slab_ll_recv_kernel", "slab_pos_kernel"],
        "nbx_multi.o": ["comm_allreduce3_kernel", "vv_pos_push_kernel"]}
KEYS = ["UBLKCP", "SYNCS", "MUFU.RSQ64H", "MUFU.RCP64H", "LDG.E.ENL2.256", "LDG.E.NA", "DFMA", "DMUL", "DADD", "DSETP", "SHFL",
        "CCTL", "ST.E", "STG", "RED", "ATOM", "BAR", "LDS", "STS"]
SHOW = ("UBLKCP", "SYNCS", "MUFU.RSQ64H", "MUFU.RCP64H", "LDG.E.ENL2.256", "LDG.E.NA")


def kernels(obj):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(O, obj)], capture_output=True, text=True).stdout
    cur, d = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if m and cur:
            d[cur].append(m.group(2).strip())
    return d


def main():
    out = ["# SASS evidence for the shipped kernels (cuobjdump -sass of the objects linked into libnbody_b200.so, sm_100a).",
           "# Regenerate: python profiles/sass_summary.py.  Static counts per instantiation (unrolled bodies, not per pair)."]
    for obj, names in WANT.items():
        for k, ins in kernels(obj).items():
            if not any(n in k for n in names):
                continue
            c = collections.Counter()
            for i in ins:
                mn = re.sub(r"^@!?U?P\d+\s+", "", i).split()[0]
                for key in KEYS:
                    if mn.startswith(key):
                        c[key] += 1
            name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:150]
            out.append(f"\n{name}\n  {len(ins)} instructions; " + ", ".join(f"{a} {b}" for a, b in c.items() if b))
            shown = set()
            for i in ins:
                for key in SHOW:
                    if key in i and key not in shown:
                        shown.add(key)
                        out.append("    e.g. " + i)
    with open(os.path.join(ROOT, "profiles", "sass_r02.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print(len(out), "lines")


if __name__ == "__main__":
    main()
