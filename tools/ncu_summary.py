"""Summarise an .ncu-rep: key raw metrics + stall hot spots per SASS instruction (needs ncu on PATH)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'lts__t_sector_hit_rate.pct',
        'launch__grid_size', 'launch__block_size', 'smsp__warps_eligible.avg.per_cycle_active']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:110])
    for w in want:
        if w in hdr:
            print("  %-78s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    vals = sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h) for h in st), reverse=True)[:8]
    for v, h in vals:
        print("  stall %7.3f %s" % (v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if hi:
    hdr = rows[hi[0]]; data = rows[hi[0] + 1: hi[1] - 1 if len(hi) > 1 else None]
    si = hdr.index('Warp Stall Sampling (All Samples)'); sc = hdr.index('Source'); ie = hdr.index('Instructions Executed')
    data = [r for r in data if len(r) > si and r[si].isdigit()]
    tot = sum(int(r[si]) for r in data); toti = sum(int(r[ie]) for r in data)
    print("samples", tot, "warp-instructions", toti, "sass lines", len(data))
    top = sorted(((int(r[si]), i, r[sc].strip(), int(r[ie])) for i, r in enumerate(data)), reverse=True)[:top_n]
    for s, i, t, e in top:
        print("%6d %5.1f%% #%4d x%-9d %s" % (s, 100 * s / tot, i, e, t[:100]))
    reg = collections.Counter()
    for i, r in enumerate(data):
        reg[i // 50] += int(r[si])
    print("by 50-instr region:", {k * 50: round(100 * v / tot, 1) for k, v in sorted(reg.items())})
    regi = collections.Counter()
    for i, r in enumerate(data):
        regi[i // 50] += int(r[ie])
    print("warp-instructions executed by 50-instr region (%):", {k * 50: round(100 * v / max(toti, 1), 1) for k, v in sorted(regi.items())})
# per CUDA source line (needs -lineinfo and --import-source on)
cu = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(cu)))
hi = [i for i, r in enumerate(rows) if r and 'Source' in r and any('Sampling' in c for c in r)]
for h in hi[:1]:
    hdr = rows[h]
    try:
        si = hdr.index('Warp Stall Sampling (All Samples)'); sc = hdr.index('Source'); ie = hdr.index('Instructions Executed')
    except ValueError:
        break
    data = [r for r in rows[h + 1:] if len(r) > max(si, sc, ie) and r[si].replace(',', '').isdigit()]
    tot = sum(int(r[si].replace(',', '')) for r in data) or 1
    print("--- by CUDA source line (samples %d)" % tot)
    top = sorted(((int(r[si].replace(',', '')), r[0], r[sc].strip(), r[ie]) for r in data), reverse=True)[:top_n]
    for s, ln, t, e in top:
        print("%6d %5.1f%% L%-5s x%-10s %s" % (s, 100 * s / tot, ln, e, t[:110]))
