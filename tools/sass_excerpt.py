"""SASS evidence for the TMA bulk copy that stages a robot's record: for every solve-kernel
instantiation, the mbarrier expect-tx arrive (SYNCS.ARRIVE.TRANS64), the bulk copy (UBLKCP.S.G) and the
try-wait loop (SYNCS.PHASECHK).  usage: python tools/sass_excerpt.py > profiles/r2_sass_tma_excerpt.txt"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "rl_mpc_locomotion_b200", "libmpc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.split("\n")
fn, ctx = None, []
print(f"# cuobjdump -sass {os.path.relpath(so, root)} | instructions around the TMA bulk copy, per kernel")
for i, l in enumerate(sass):
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = m.group(1)
    if "UBLKCP" in l:
        dm = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip()
        mm = re.search(r"(mpc_\w+<[^>]*>)", dm)
        name = mm.group(1) if mm else dm[:80]
        print(f"\n## {name}")
        for k in range(max(0, i - 8), min(len(sass), i + 3)):
            if re.search(r"SYNCS|UBLKCP|FENCE|UMOV|S2UR", sass[k]):
                print(sass[k].rstrip()[:120])
        for k in range(i, min(len(sass), i + 400)):   # the wait loop
            if "SYNCS.PHASECHK" in sass[k]:
                print(sass[k].rstrip()[:120])
                break
