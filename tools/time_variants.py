# times the c2 frame (device-resident film) with every experiment build under tools/_variants (one subprocess per library)
import glob, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = sys.argv[1:] or sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(root, "tools/_variants/*.so")))
for rep in range(2):
    for n in names:
        env = dict(os.environ, VDBRT_LIBRARY=os.path.join(root, "tools/_variants", n + ".so"))
        r = subprocess.run([sys.executable, os.path.join(root, "tools/prof_c2.py"), "8", "device"], env=env, capture_output=True, text=True)
        ms = [float(l.strip("()\n").split(",")[0]) for l in r.stdout.splitlines() if l.startswith("(")]
        print("%-24s min %.3f  med %.3f ms" % (n, min(ms), sorted(ms)[len(ms) // 2]) if ms else (n, r.stderr[-500:]))
