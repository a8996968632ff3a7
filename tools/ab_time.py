"""A/B device timing of several builds of the library in one process-per-build loop.
usage: python tools/ab_time.py libA.so libB.so ...   (prints the OSQP-faithful and fixed-1 times)"""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rep in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, MPC_B200_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "quick_time.py"), "4096", "10"], env=env,
                             capture_output=True, text=True).stdout
        keep = [l.split(":")[1].split("->")[0].strip() for l in out.split("\n") if " osqp   " in l or "fixed105" in l or "fixed1_noscale" in l]
        print(f"{os.path.basename(lib):28s} osqp / fixed105 / fixed1_noscale: {' / '.join(keep)}", flush=True)
