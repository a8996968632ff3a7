"""Where the local-memory (spill) instructions of the h = 10 solve kernel sit, by source line (needs -lineinfo).
usage: python tools/sass_spills.py [lib.so]"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "rl_mpc_locomotion_b200", "libmpc_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text.") and (sys.argv[2] if len(sys.argv) > 2 else "mpc_solve_kernelILi10ELb0ELi60ELb0") in l][0]
cur, cnt, tot, n = None, collections.Counter(), collections.Counter(), 0
for l in txt[start + 1:]:
    if (l.startswith("//---------------------") or l.startswith("\t.section")) and n:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        n += 1
        tot[cur] += 1
        if re.search(r"\b(LDL|STL)\b", m.group(2)):
            cnt[cur] += 1
print(f"total SASS instructions {n}, local-memory instructions {sum(cnt.values())}")
for k, c in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{k[0]}:{k[1]:5d}  {c:4d} LDL/STL of {tot[k]} instr")
