"""One markdown table row per distinct kernel of an .ncu-rep (the longest launch of each): duration, DRAM bytes and
throughput, occupancy limits, registers, shared memory, issue activity, the two largest stall reasons.
usage: python tools/ncu_table.py rep.ncu-rep [rep2 ...]"""
import csv, io, re, subprocess, sys

COLS = {"t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "regs": "launch__registers_per_thread", "smem": "launch__shared_mem_per_block_dynamic", "smem_s": "launch__shared_mem_per_block_static",
        "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "grid": "launch__grid_size", "block": "launch__block_size",
        "l1": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "occ": "launch__occupancy_limit_shared_mem", "occr": "launch__occupancy_limit_registers"}
UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = re.sub(r"\(.*", "", name).replace("lsc::", "").replace("void ", "")
    name = re.sub(r"unsigned long", "u64", name); name = re.sub(r"unsigned int", "u32", name); name = re.sub(r"unsigned char", "u8", name)
    return name.replace("(int)", "").replace("(bool)", "")[:90]


best = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ix = {k: hdr.index(v) for k, v in COLS.items() if v in hdr}
    stall = [(i, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
             if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for r in rows[2:]:
        def val(k, default=0.0):
            if k not in ix:
                return default
            try:
                return float(r[ix[k]].replace(",", "")) * UNIT.get(units[ix[k]], 1.0)
            except ValueError:
                return default
        name = short(r[hdr.index("Kernel Name")])
        t = val("t")
        if name in best and best[name]["t"] >= t:
            continue
        st = sorted(((float(r[i].replace(",", "") or 0), h) for i, h in stall), reverse=True)[:2]
        best[name] = {"t": t, "gb": (val("rd") + val("wr")) / 1e9, "dram": val("dram"), "warps": val("warps"), "regs": int(val("regs")),
                      "smem": (val("smem") + val("smem_s")) / 1e3, "issue": val("issue"), "grid": int(val("grid")), "block": int(val("block")),
                      "lim": "smem %d / regs %d" % (int(val("occ")), int(val("occr"))), "st": ", ".join("%s %.1f" % (h, v) for v, h in st)}
print("| kernel | ms | DRAM GB (TB/s) | DRAM % of ncu peak | warps active % | issue active % | regs | smem KB | CTAs/SM limit | grid x block | top stalls (per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for name, b in sorted(best.items(), key=lambda kv: -kv[1]["t"]):
    print("| `%s` | %.3f | %.2f (%.2f) | %.0f | %.0f | %.0f | %d | %.1f | %s | %d x %d | %s |" % (
        name, b["t"], b["gb"], b["gb"] / max(b["t"], 1e-9), b["dram"], b["warps"], b["issue"], b["regs"], b["smem"], b["lim"], b["grid"], b["block"], b["st"]))
