"""Aggregate an ncu report's per-SASS metrics by CUDA source line (needs -lineinfo).
usage: python tools/ncu_by_line.py gpurun_out/prof.ncu-rep [H] [topN]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep = sys.argv[1]
H = sys.argv[2] if len(sys.argv) > 2 else "10"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("NCU_LIB", os.path.join(root, "rl_mpc_locomotion_b200", "libmpc_b200.so"))
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
STALLS = ["stall_no_inst", "stall_barrier", "stall_short_sb", "stall_long_sb", "stall_mio", "stall_wait", "stall_math"]
ist = [hdr.index(k) for k in STALLS]
assert len(seq) == len(data), (len(seq), len(data))
agg = collections.defaultdict(lambda: [0, 0] + [0] * len(STALLS))
tot_st = [0] * len(STALLS)
ti = ts = 0
for (loc, ins), r in zip(seq, data):
    a, b = int(r[ii]), int(r[isamp])
    agg[loc][0] += a; agg[loc][1] += b; ti += a; ts += b
    for q, ix in enumerate(ist):
        v = int(r[ix] or 0); agg[loc][2 + q] += v; tot_st[q] += v
src = {}
print(f"SASS instructions {len(seq)}, executed {ti}, samples {ts}")
print("stall samples: " + ", ".join(f"{k[6:]} {v / ts * 100:.1f}%" for k, v in zip(STALLS, tot_st)))
# coarse regions (50-line windows of the solver source) by samples
reg = collections.defaultdict(lambda: [0, 0, 0, 0])
for loc, v in agg.items():
    key = (loc[0], loc[1] // 50 * 50) if loc else None
    reg[key][0] += v[0]; reg[key][1] += v[1]; reg[key][2] += v[2]; reg[key][3] += v[3]
print("regions (50 source lines): inst% samp% no_inst% barrier%")
for key, v in sorted(reg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {str(key):34s} {v[0] / ti * 100:5.1f} {v[1] / ts * 100:5.1f} {v[2] / ts * 100:5.1f} {v[3] / ts * 100:5.1f}")
for loc, (a, b, *st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    line = ""
    if loc:
        fn = os.path.join(root, "rl_mpc_locomotion_b200", "csrc", loc[0])
        if os.path.exists(fn):
            src.setdefault(fn, open(fn).read().split("\n"))
            line = src[fn][loc[1] - 1].strip()[:80]
    print(f"{str(loc):30s} inst {a / ti * 100:5.1f}% samp {b / ts * 100:5.1f}% noinst {st[0] / ts * 100:4.1f}% bar {st[1] / ts * 100:4.1f}%  {line}")
