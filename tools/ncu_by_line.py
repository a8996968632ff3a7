"""Aggregate an ncu report's per-SASS metrics by CUDA source line (needs -lineinfo).
usage: python tools/ncu_by_line.py gpurun_out/prof.ncu-rep [H] [topN]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep = sys.argv[1]
H = sys.argv[2] if len(sys.argv) > 2 else "10"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "rl_mpc_locomotion_b200", "libmpc_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text.") and f"mpc_solve_kernelILi{H}ELb0ELi{6*int(H)}ELb0" in l][0]
cur, seq = None, []
for l in txt[start + 1:]:
    if (l.startswith("//---------------------") or l.startswith("\t.section")) and seq:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(seq) == len(data), (len(seq), len(data))
agg = collections.defaultdict(lambda: [0, 0])
ti = ts = 0
for (loc, ins), r in zip(seq, data):
    a, b = int(r[ii]), int(r[isamp])
    agg[loc][0] += a; agg[loc][1] += b; ti += a; ts += b
src = {}
print(f"SASS instructions {len(seq)}, executed {ti}, samples {ts}")
for loc, (a, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    line = ""
    if loc:
        fn = os.path.join(root, "rl_mpc_locomotion_b200", "csrc", loc[0])
        if os.path.exists(fn):
            src.setdefault(fn, open(fn).read().split("\n"))
            line = src[fn][loc[1] - 1].strip()[:80]
    print(f"{str(loc):30s} inst {a / ti * 100:5.1f}% samp {b / ts * 100:5.1f}%  {line}")
