#!/usr/bin/env python3
"""Join the per-SASS-instruction counters of an ncu report (--set full --import-source on) with the line table of the
library it profiled: dynamic warp instructions, active lanes and stall samples per source line / per code group.
usage: tools/ncu_lines.py REPORT.ncu-rep LIBRARY.so [kernel-substring] [--top N]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

def sass_lines(lib, kern):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    out = []
    for f in os.listdir(d):
        if not f.endswith(".cubin"): continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        for fn in re.split(r'\n\s*\.text\.', txt)[1:]:
            if kern not in fn.split('\n')[0]: continue
            cur = None
            for line in fn.split('\n'):
                m = re.search(r'//## File "([^"]+)", line (\d+)', line)
                if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
                m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
                if m: out.append((int(m.group(1), 16), cur, m.group(2)))
            return out
    return out

def main():
    rep, lib = sys.argv[1], sys.argv[2]
    kern = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "k_render_levelsetILb0ELb0"
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    lines = sass_lines(lib, kern)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = (rows[1], rows[2:]) if "Address" in rows[1] else (rows[0], rows[1:])
    iA, iE, iT, iS = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
    assert len(data) == len(lines), (len(data), len(lines))
    per = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
    for k, r in enumerate(data):
        cur = lines[k][1]
        for j, i in enumerate((iE, iT, iS)):
            per[cur][j] += int(r[i]); tot[j] += int(r[i])
    print("SASS %d instr; executed %.1fM warp instr, %.2f lanes avg, %d samples" % (len(lines), tot[0] / 1e6, tot[1] / max(tot[0], 1), tot[2]))
    byfile = collections.defaultdict(list)
    for k, v in per.items():
        if k: byfile[k[0]].append((k[1], v))
    for cur, v in sorted(per.items(), key=lambda x: -x[1][0])[:top]:
        print("%-28s %8.1fM %5.1f%%  lanes %5.1f  samples %5.1f%%" % ("%s:%d" % cur if cur else "?", v[0] / 1e6, 100 * v[0] / tot[0], v[1] / max(v[0], 1), 100 * v[2] / max(tot[2], 1)))
    # per function-ish group: contiguous line ranges of vdbrt_device.cuh given on the command line as --groups a-b,c-d
    if "--groups" in sys.argv:
        for g in sys.argv[sys.argv.index("--groups") + 1].split(","):
            f, r = g.split(":"); a, b = map(int, r.split("-"))
            e = sum(v[0] for k, v in per.items() if k and k[0].startswith(f) and a <= k[1] <= b)
            t = sum(v[1] for k, v in per.items() if k and k[0].startswith(f) and a <= k[1] <= b)
            s = sum(v[2] for k, v in per.items() if k and k[0].startswith(f) and a <= k[1] <= b)
            print("group %-26s %8.1fM %5.1f%% lanes %5.1f samples %5.1f%%" % (g, e / 1e6, 100 * e / tot[0], t / max(e, 1), 100 * s / max(tot[2], 1)))

if __name__ == "__main__":
    main()
