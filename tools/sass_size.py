"""Static SASS size of the solve kernel by source region (needs -lineinfo): where the instruction-
cache footprint comes from.  usage: python tools/sass_size.py [H] [lib.so]"""
import collections, os, re, subprocess, sys, tempfile

H = sys.argv[1] if len(sys.argv) > 1 else "10"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "rl_mpc_locomotion_b200", "libmpc_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text.") and f"mpc_solve_kernelILi{H}ELb0ELi{6*int(H)}ELb0" in l][0]
cur, cnt, n = None, collections.Counter(), 0
for l in txt[start + 1:]:
    if (l.startswith("//---------------------") or l.startswith("\t.section")) and n:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l):
        cnt[cur] += 1; n += 1
print(f"total SASS instructions {n} = {n * 16 / 1024:.0f} KiB")
# bucket by 25-line windows of the solver source
b = collections.Counter()
for (f, ln), c in cnt.items():
    b[(f, ln // 5 * 5)] += c
for (f, ln), c in sorted(b.items(), key=lambda kv: -kv[1])[:200]:
    print(f"{f}:{ln:5d}-{ln+4:5d}  {c:6d} instr")
