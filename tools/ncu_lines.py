"""Stall samples and executed instructions per CUDA source line of one profiled launch of an .ncu-rep
(needs -lineinfo at compile time and --import-source on at capture time).
usage: python tools/ncu_lines.py rep.ncu-rep <launch index> [top_n]"""
import csv, io, subprocess, sys, collections
rep, idx = sys.argv[1], int(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-s", str(idx), "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict()
fname, func, hdr = "", "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        func = r[1]; continue
    if r[0] == "Line No":
        hdr = r; si = hdr.index("Warp Stall Sampling (All Samples)"); ie = hdr.index("Instructions Executed"); continue
    if hdr and r[0].isdigit() and len(r) > max(si, ie):
        try:
            s = int(r[si]); e = int(r[ie])
        except ValueError:
            continue
        k = (fname, int(r[0]))
        a = agg.setdefault(k, [0, 0, r[1].strip()])
        a[0] += s; a[1] += e
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print(func[:150])
print("samples %d warp-instructions %d" % (tot, toti))
for (f, ln), (s, e, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print("%6d %5.1f%% | instr %5.1f%% | %s:%d  %s" % (s, 100.0 * s / tot, 100.0 * e / toti, f, ln, src[:100]))
