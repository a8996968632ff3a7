"""Record what one `ncu --set full` capture says about the dominant kernel of a bench workload in
profiles/dram_traffic.json, stamped with the hash of the kernel sources it was taken on (bench.py quotes the
figures only while that hash matches its own build).
usage: python tools/ncu_to_traffic.py <workload> <x.ncu-rep> <profiles/summary.txt it is summarised in>"""
import csv, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import bench
workload, rep, src = sys.argv[1:4]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m, unit = dict(zip(hdr, vals)), dict(zip(hdr, units))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
f = lambda k: float(m[k].replace(",", "")) * scale.get(unit[k], 1.0)
rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
wave, cyc = f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), f("sm__cycles_elapsed.max")
path = os.path.join(root, "profiles", "dram_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[workload] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": m.get("Kernel Name"),
               "source": src, "kernel_hash": bench.kernel_hash(),
               "binding_roof": "shared-memory bandwidth (l1tex lsu shared wavefronts, 1 per cycle per SM)",
               "binding_frac_of_peak": wave / (148 * cyc), "lsu_shared_wavefronts": wave, "sm_cycles_elapsed_max": cyc}
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(d[workload], indent=1))
