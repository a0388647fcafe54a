#!/usr/bin/env python
"""bench.py -- headline benchmark of libsais_cuda (contract: see the task statement / DESIGN.md §6).

N = 1  (python bench.py): BASELINE.json configs[1] -- libsais_bwt + primary index of a 256 MiB synthetic
  random-byte text (splitmix64 generator, seed 2), one text per step.
    value : MB/s of input, device-timed with CUDA events on the context's stream, input and output resident
            in HBM (libsais_cuda_bwt_dev).  Per-kernel CUDA events stay on inside the timed region (the
            roofline figures are measured there); their cost is ~50 event records per 8 ms step.
    e2e   : same metric through the reference-facing C-ABI call libsais_bwt_ctx() with HOST (pinned) buffers:
            H2D of the text and D2H of the BWT inside the timed region.
    roofline : the dominant kernel (largest share of the step), algorithmic bytes / CUDA-event time measured
            live inside the timed region against MEASURED_PEAKS.json's hbm_gbs; plus `whole_step` (all bytes
            moved by all kernels / step time) and `worst_kernel` (lowest fraction among kernels >= 3 % of the step).
    cpu_baseline : the unmodified reference (oracle/_ref, libsais_bwt_omp) on this box's host cores on the FULL
            256 MiB text, best of a thread sweep; its BWT and primary index are compared byte for byte with
            the GPU's.
    c3, c4 : sub-records -- configs[2] (libsais + libsais_plcp + libsais_lcp on 1.9 GB repetitive DNA,
            generated on the device, verified with the linear-time checker) and configs[3] on one GPU
            (64 x 128 MiB DNA blocks through libsais_cuda_bwt_batch).
N > 1 (torchrun): BASELINE.json configs[3] -- the fixed batch of 64 x 128 MiB independent DNA blocks, block b on
  rank b mod N, no collective on the data path -> "strong" scaling.  value: device-timed (blocks resident in HBM);
  e2e: libsais_cuda_bwt_batch on pinned host buffers (3 host threads / contexts per GPU overlap copies and kernels).
`--impl reference` times the reference's own CPU implementation on the same config (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FULL = 1 << 28            # 256 MiB (configs[1])
SEED = 2
C4_BLOCKS = 64              # configs[3]
C4_BLOCK_BYTES = 1 << 27
C3_BASE, C3_COPIES = 19_000_000, 100
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
WORKLOAD_C2 = "libsais_bwt + primary index, 256 MiB iid random bytes per step (BASELINE configs[1]; splitmix64 seed 2)"
WORKLOAD_C4 = "batch of 64 x 128 MiB independent libsais_bwt blocks (iid ACGT, seed 1000+b), block b on GPU b mod N (BASELINE configs[3])"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons; samples are kept only if they fall inside the timed
    region [mark_start(), mark_stop()] (nvidia-smi needs a moment to start, so it is launched early)."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc, self.t0, self.t1 = [], None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, mx, reasons, allsm = [], 0, set(), []
        for ts, l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                v = float(f[1]); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            allsm.append(v)
            if self.t0 is not None and not (self.t0 - 0.02 <= ts <= (self.t1 or ts) + 0.05):
                continue
            sm.append(v)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        use = sm if sm else allsm
        return {"sm_mhz": float(np.median(use)) if use else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_total": len(allsm)}


# --------------------------------------------------------------------------------------------- reference (CPU)
def _cpu_flags():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


_ISA_PREFIXES = ("avx", "amx", "sse", "ssse", "bmi", "fma", "f16c", "gfni", "vaes", "vpclmul", "sha", "movdir", "cldemote",
                 "serialize", "adx", "rdseed", "popcnt", "movbe", "abm", "pclmul", "aes", "lzcnt", "clwb", "clflushopt", "rdpid", "waitpkg")


def ref_lib():
    """The unmodified reference: the -march=native build when this box's CPU has every ISA flag of the build host
    (the author's recipe, reference Benchmarks.md:5), else the portable -march=x86-64-v3 build."""
    d = os.path.join(ROOT, "oracle", "_ref")
    portable, native = os.path.join(d, "libsais_ref.so"), os.path.join(d, "libsais_ref_native.so")
    if not os.path.exists(portable):
        if os.path.isdir("/root/reference/src"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
        else:
            return None, None
    path, build = portable, "gcc -O3 -march=x86-64-v3 -fopenmp"
    try:
        if os.path.exists(native):
            need = {f for f in open(os.path.join(d, "native_flags.txt")).read().split() if f.startswith(_ISA_PREFIXES)}
            if need and need <= _cpu_flags():
                march = open(os.path.join(d, "native_march.txt")).read().strip()
                path, build = native, "gcc -O3 -march=native (=%s, ISA flags verified on this host) -fopenmp" % march
    except OSError:
        pass
    lib = C.CDLL(path)
    lib.libsais_bwt_omp.restype = C.c_int32
    return lib, build


def cpu_bwt(lib, T, threads):
    """One reference libsais_bwt_omp call: (seconds, primary index, U)."""
    n = len(T)
    U = np.empty(n, dtype=np.uint8)
    A = np.empty(n, dtype=np.int32)
    t0 = time.perf_counter()
    rc = lib.libsais_bwt_omp(T.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p),
                             C.c_int32(n), C.c_int32(0), None, C.c_int32(threads))
    dt = time.perf_counter() - t0
    assert rc > 0, "reference libsais_bwt_omp returned %d" % rc
    return dt, rc, U


def thread_candidates():
    ncpu = os.cpu_count() or 1
    return sorted({min(8, ncpu), min(16, ncpu), min(32, ncpu), ncpu})


def cpu_baseline_c2(T, gpu_U, gpu_primary):
    """Reference libsais_bwt_omp on the FULL text, one call per thread count; byte-for-byte parity with the GPU."""
    lib, build = ref_lib()
    if lib is None:
        return None
    best, parity = None, True
    sweep = {}
    for th in thread_candidates():
        dt, rc, U = cpu_bwt(lib, T, th)
        sweep[str(th)] = round(len(T) / 1e6 / dt, 2)
        parity = parity and rc == gpu_primary and bool(np.array_equal(U, gpu_U))
        if best is None or dt < best[0]:
            best = (dt, th)
    return {"value": round(len(T) / 1e6 / best[0], 2), "unit": "MB/s", "cores": best[1], "kind": "reference",
            "sample": "libsais_bwt_omp (libsais 2.10.4, %s) on the FULL %d MiB text of this run, one call per thread count %s, best quoted; "
                      "host has %d logical CPUs" % (build, len(T) >> 20, sorted(int(k) for k in sweep), os.cpu_count() or 1),
            "threads_sweep_mbs": sweep, "gpu_bwt_equals_reference": parity}


def run_reference_arm(args):
    """The reference's own CPU implementation on this arm's config.  N = 1: the full 256 MiB text per step.
    N > 1: one whole 128 MiB block of the 64-block batch per step (blocks are independent and equally sized, so a
    step is a bounded sample of the batch at full block size)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from libsais_b200 import gen
    lib, build = ref_lib()
    ncpu = os.cpu_count() or 1
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsais_ref.so missing and /root/reference absent"}))
        return
    c4 = args.gpus > 1
    if c4:
        n = C4_BLOCK_BYTES
        texts = [gen.dna(1000 + b, n) for b in range(min(2, max(1, args.steps)))]
        workload = WORKLOAD_C4
        sample = "each step = libsais_bwt_omp on ONE full 128 MiB block of the 64-block batch (block = step mod %d)" % len(texts)
    else:
        n = args.n
        texts = [gen.rand_bytes(SEED, n)]
        workload = WORKLOAD_C2 if n == N_FULL else "libsais_bwt + primary index, %d MiB iid random bytes per step" % (n >> 20)
        sample = "each step = libsais_bwt_omp on the full %d MiB text" % (n >> 20)
    # thread sweep (also the warm-up): one full-size call per candidate
    sweep, best = {}, None
    for th in thread_candidates():
        dt, _, _ = cpu_bwt(lib, texts[0], th)
        sweep[str(th)] = round(n / 1e6 / dt, 2)
        if best is None or dt < best[0]:
            best = (dt, th)
    threads = best[1]
    for _ in range(max(0, args.warmup - len(sweep))):
        cpu_bwt(lib, texts[0], threads)
    # keep the whole arm within a few minutes: at most ~240 s of timed CPU work
    steps_timed = max(1, min(args.steps, int(240.0 / max(best[0], 1e-3))))
    t0 = time.perf_counter()
    for s in range(steps_timed):
        cpu_bwt(lib, texts[s % len(texts)], threads)
    dt = time.perf_counter() - t0
    v = n * steps_timed / 1e6 / dt
    print(json.dumps({
        "impl": "reference", "metric": "bwt_construction_throughput", "value": round(v, 2), "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "steps_timed": steps_timed, "warmup": args.warmup, "ms_per_step": round(dt / steps_timed * 1e3, 2),
        "higher_is_better": True, "scaling": "strong" if c4 else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "text_bytes": n},
        "cpu_baseline": {"value": round(v, 2), "unit": "MB/s", "cores": threads, "kind": "reference",
                         "sample": sample + "; %s; best of thread sweep %s; host has %d logical CPUs" % (build, sweep, ncpu)},
        "e2e": {"value": round(v, 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------------------------- helpers (GPU arm)
def numa_bind(local):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so pinned host buffers are allocated
    next to the GPU's PCIe root (8 ranks sharing one node was the end-to-end limiter at N = 8 in round 1)."""
    try:
        import torch
        props = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "numa_node=-1"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "node %d (%d cpus)" % (node, len(allowed))
        return "node %d (no allowed cpus)" % node
    except Exception as e:                                   # topology hidden (VM) or no permission: run unbound
        return "unbound (%s)" % type(e).__name__


def host_link_probe(dev, barrier, dist, nbytes=1 << 28, reps=4):
    """What the host <-> device links give when every rank copies at once: pinned H2D and D2H of `nbytes`, concurrently on two
    streams, no kernels.  The end-to-end arm cannot beat this floor; on a shared / virtualised host it is well below 8 x PCIe."""
    import torch
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev); d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    def once():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    once(); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    world = dist.get_world_size() if dist is not None else 1
    gbs = world * reps * nbytes / float(t[0]) / 1e9
    del h_in, h_out, d_in, d_out
    return {"h2d_plus_d2h_concurrent_gbs_each_direction_all_ranks": round(gbs, 1), "per_rank_gbs_each_direction": round(gbs / world, 1),
            "bytes_per_copy": nbytes}


def kernel_tables(agg, steps, dev_ms, peak):
    kern, fr = {}, []
    tot_bytes = 0.0
    for k, v in agg.items():
        gbs = v["bytes"] / max(v["ms"], 1e-9) / 1e6
        kern[k] = {"launches": v["launches"], "ms_per_step": round(v["ms"] / steps, 4), "algo_gbs": round(gbs, 1),
                   "frac": round(gbs / peak, 4), "share_of_step": round(v["ms"] / dev_ms, 4)}
        tot_bytes += v["bytes"]
        if v["ms"] / dev_ms >= 0.03 and v["bytes"] > 0:
            fr.append((gbs / peak, k))
    return kern, tot_bytes, (min(fr) if fr else None)


TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic_r2.json")


def roofline_record(agg, steps, dev_ms, peak, peak_src, kernel_names):
    kern, tot_bytes, worst = kernel_tables(agg, steps, dev_ms, peak)
    dom_name = max(agg, key=lambda k: agg[k]["ms"])
    dom = agg[dom_name]
    achieved = dom["bytes"] / (dom["ms"] / 1e3) / 1e9
    traffic = None
    if os.path.exists(TRAFFIC_FILE):
        try:
            traffic = json.load(open(TRAFFIC_FILE)).get(dom_name, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
    whole = tot_bytes / (dev_ms / 1e3) / 1e9
    roof = {"bound": "hbm", "kernel": kernel_names.get(dom_name, dom_name), "kernel_class": dom_name, "achieved": round(achieved, 1), "peak": peak,
            "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
            "algorithmic_bytes_per_launch": dom["bytes"] / max(dom["launches"], 1),
            "avg_launch_ms": dom["ms"] / max(dom["launches"], 1), "launches": dom["launches"], "share_of_step": round(dom["ms"] / dev_ms, 4),
            "whole_step": {"bytes_moved_per_step": tot_bytes / steps, "achieved": round(whole, 1), "frac": round(whole / peak, 4),
                           "note": "sum of the algorithmic bytes of every kernel launched in the step / device time of the step"},
            "worst_kernel": None if worst is None else {"kernel_class": worst[1], "frac": round(worst[0], 4),
                                                        "share_of_step": kern[worst[1]]["share_of_step"]},
            "note": "achieved = sum of algorithmic bytes / sum of CUDA-event time over all launches of the kernel class in the timed region; "
                    "traffic = dram read+write bytes of one launch of it from the committed ncu capture (profiles/)"}
    return roof, kern


KERNEL_NAMES = {
    "sort_pass": "sort_pass_kernel<u64,u32> (stable onesweep digit pass)",
    "sort_pass_gen": "part_pass_kernel<KmerSrc> (first MSD level: keys built from the packed text, partition by the top 8 key bits)",
    "part_pass": "part_pass_kernel<ArraySrc,segmented> (second MSD level: unstable partition by key bits 8..15)",
    "bucket_sort": "bucket_sort_kernel<k32,fused> (16-bit buckets finished in shared memory; emits positions, BWT rows, head/active flags)",
    "rank_init": "rank_agg (MSD path) or rank_flags, rank_apply<round 0>",
    "local_sort": "po_round_kernel (rounds >= 1 on the position-ordered list; slot-ordered fallback: local_count/local_sort)",
    "round_keys": "group reorder of the position-ordered rounds (table, sort of the groups, scan, move) / round_keys",
    "scatter": "po_apply_kernel (ISA updates of a round) / partitioned scatter (round-0 ranks, phi)", "rank_update": "rank_flags/rank_apply<rounds >= 1>",
}


def accumulate(agg, st):
    for k, v in st["kernels"].items():
        a = agg.setdefault(k, {"launches": 0, "ms": 0.0, "bytes": 0.0})
        a["launches"] += v["launches"]; a["ms"] += v["ms"]; a["bytes"] += v["bytes"]
    return st["total_launches"]


# --------------------------------------------------------------------------------------------- config 3 (sub-record, N = 1)
def run_c3(ctx, dev, peak):
    import torch
    from libsais_b200 import gen, check
    out = {"workload": "libsais + libsais_plcp + libsais_lcp, 1.9e9 B repetitive DNA (19 M base x 100 copies, 1e-3 substitutions; BASELINE configs[2]), generated on the device"}
    n = C3_BASE * C3_COPIES
    free, _ = torch.cuda.mem_get_info()
    if free < 140e9:
        out["skipped"] = "needs ~125 GB of free HBM, %.0f GB free" % (free / 1e9)
        return out
    t0 = time.time()
    dT = gen.repetitive_dna_torch(C3_BASE, C3_COPIES, device=dev)
    torch.cuda.synchronize()
    out["n"] = n; out["gen_s"] = round(time.time() - t0, 1)
    dSA = torch.empty(n, dtype=torch.int32, device=dev)
    ctx.set_profiling(True)

    def timed(fn, reps=2):
        best, st = None, None
        for _ in range(reps + 1):                        # first call warms the workspace
            rc = fn()
            s = ctx.stats()
            if best is None or s["device_ms"] < best:
                best, st = s["device_ms"], s
        return rc, best, st

    rc, ms, st = timed(lambda: ctx.sa_dev(dT.data_ptr(), dSA.data_ptr(), n), reps=1)
    agg = {}
    accumulate(agg, st)
    kern, tot_bytes, worst = kernel_tables(agg, 1, ms, peak)
    out["sa"] = {"rc": rc, "ms": round(ms, 2), "mbs": round(n / 1e6 / (ms / 1e3), 1), "launches": st["total_launches"],
                 "rounds": [(r["h"], r["n_active"], r["passes"]) for r in st["rounds"]],
                 "per_round": [{"h": r["h"], "n_active": r["n_active"], "ms": round(r["ms"], 2), "bytes_moved": r["bytes"],
                                "frac_of_peak": round(r["bytes"] / max(r["ms"], 1e-9) / 1e6 / peak, 4)} for r in st["rounds"]],
                 "bytes_moved": tot_bytes, "moved_bytes_frac_of_peak": round(tot_bytes / (ms / 1e3) / 1e9 / peak, 4),
                 "kernels": {k: v for k, v in kern.items() if v["share_of_step"] >= 0.01}}
    out["sa"]["verify"] = check.verify_sa(dT, dSA, n)
    ctx.release_workspace()
    torch.cuda.empty_cache()
    dP = torch.empty(n, dtype=torch.int32, device=dev)
    rc, ms, st = timed(lambda: ctx.plcp_dev(dT.data_ptr(), dSA.data_ptr(), dP.data_ptr(), n))
    agg = {}; accumulate(agg, st)
    kern, tot_bytes, _ = kernel_tables(agg, 1, ms, peak)
    out["plcp"] = {"rc": rc, "ms": round(ms, 2), "mbs": round(n / 1e6 / (ms / 1e3), 1), "max": int(dP.max()), "mean": round(float(dP.double().mean()), 1),
                   "moved_bytes_frac_of_peak": round(tot_bytes / (ms / 1e3) / 1e9 / peak, 4), "kernels": kern}
    T_host = dT.cpu().numpy()
    out["plcp"]["verify_sample"] = check.verify_plcp_sample(T_host, dSA, dP, n, samples=5000)
    del T_host
    ctx.release_workspace()
    dL = torch.empty(n, dtype=torch.int32, device=dev)
    rc, ms, st = timed(lambda: ctx.lcp_dev(dP.data_ptr(), dSA.data_ptr(), dL.data_ptr(), n))
    out["lcp"] = {"rc": rc, "ms": round(ms, 2), "mbs": round(n / 1e6 / (ms / 1e3), 1),
                  "moved_bytes_frac_of_peak": round(12.0 * n / (ms / 1e3) / 1e9 / peak, 4), "verify": check.verify_lcp(dP, dSA, dL, n)}
    del dT, dSA, dP, dL
    ctx.release_workspace()
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------- config 4 (batch of blocks)
def run_c4(args, ctx, dev, local, rank, world, dist, barrier, sampler=None, reps=2, warmup=0):
    """The 64-block batch: this rank's blocks are b = rank, rank + world, ...  Returns timings (max over ranks)."""
    import torch
    import libsais_b200
    from libsais_b200 import gen
    lib = libsais_b200.load_library()
    nblk, n = args.c4_blocks, args.c4_block_bytes
    from libsais_b200 import sharding
    mine = sharding.blocks_for_rank(nblk, rank, world)          # block b -> rank b mod world (tests/test_sharding.py covers it on gloo)
    k = len(mine)
    # inputs: generated on the device (numpy needs ~2 s per block), resident copy + pinned host copy
    dT = torch.empty((k, n), dtype=torch.uint8, device=dev)
    for i, b in enumerate(mine):
        dT[i] = gen.dna_torch(1000 + b, n, device=dev)
    dU = torch.empty((k, n), dtype=torch.uint8, device=dev)
    hT = torch.empty((k, n), dtype=torch.uint8).pin_memory()
    hU = torch.empty((k, n), dtype=torch.uint8).pin_memory()
    hT.copy_(dT)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    # ---- device-resident: one context, block after block
    ctx.set_profiling(True)
    prim = [0] * k
    for _ in range(max(1, warmup)):                       # warm-up passes over the whole batch (sub-record: two blocks)
        for i in range(k if warmup else min(k, 2)):
            prim[i] = ctx.bwt_dev(dT[i].data_ptr(), dU[i].data_ptr(), n)
    barrier()
    if sampler:
        sampler.mark_start()
    agg, launches = {}, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        for i in range(k):
            prim[i] = ctx.bwt_dev(dT[i].data_ptr(), dU[i].data_ptr(), n)
            assert prim[i] > 0, "bwt_dev failed on block %d: %d" % (mine[i], prim[i])
            launches += accumulate(agg, ctx.stats())
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1) / reps
    if sampler:
        sampler.mark_stop()
    rounds = ctx.stats()["rounds"]
    # ---- end to end: the batch entry point on pinned host buffers
    ctx.set_profiling(False)
    ctx.release_workspace()
    Tp = (C.c_void_p * k)(*[hT[i].data_ptr() for i in range(k)])
    Up = (C.c_void_p * k)(*[hU[i].data_ptr() for i in range(k)])
    ns = (C.c_int32 * k)(*([n] * k))
    pr = (C.c_int32 * k)()
    devs = (C.c_int32 * 1)(local)
    lib.libsais_cuda_bwt_batch.restype = C.c_int32

    def batch(cnt):
        return lib.libsais_cuda_bwt_batch(Tp, Up, ns, pr, None, C.c_int32(cnt), devs, C.c_int32(1), C.c_int32(args.lanes))


    lanes_tried = {}
    e2e_s = None
    for lanes in sorted({int(x) for x in str(args.lanes_sweep).split(',') if x} | {args.lanes}):
        args_lanes_saved = args.lanes
        args.lanes = lanes
        assert batch(min(k, 2 * max(lanes, 1))) == 0             # warm the pooled contexts' workspaces
        barrier()
        e2e_reps = min(reps, 5)                            # the device-timed arm runs exactly `reps` passes; this one at most 5 per lane count
        t0 = time.perf_counter()
        for _ in range(e2e_reps):
            rc = batch(k)
        dt = (time.perf_counter() - t0) / e2e_reps
        assert rc == 0, "libsais_cuda_bwt_batch failed"
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        lanes_tried[lanes] = float(tt[0])
        args.lanes = args_lanes_saved
    best_lanes = min(lanes_tried, key=lanes_tried.get)
    e2e_s = lanes_tried[best_lanes]
    probe = host_link_probe(dev, barrier, dist)
    # parity: host-API result == device-API result for every block, and every block inverts back to its text
    ok = all(int(pr[i]) == prim[i] for i in range(k)) and bool(torch.equal(hU.to(dev), dU))
    dB = torch.empty(n, dtype=torch.uint8, device=dev)
    inv = True
    for i in range(k):
        inv = inv and ctx.unbwt_dev(dU[i].data_ptr(), dB.data_ptr(), n, prim[i]) == 0 and bool(torch.equal(dB, dT[i]))
    lib.libsais_cuda_batch_release()
    times = torch.tensor([dev_ms, 0.0, 0.0 if (ok and inv) else 1.0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    del dT, dU, hT, hU, dB
    torch.cuda.empty_cache()
    return {"dev_ms": float(times[0]), "e2e_ms": e2e_s * 1e3, "parity_ok": float(times[2]) == 0.0, "agg": agg, "launches": launches,
            "rounds": rounds, "blocks": nblk, "block_bytes": n, "blocks_this_rank": k, "reps": reps, "probe": probe,
            "lanes": best_lanes, "lanes_tried_ms": {str(a): round(b * 1e3, 2) for a, b in lanes_tried.items()}}


def e2e_floor(r):
    """Host-link floor of the batch's end-to-end time: all input bytes in and all output bytes out at the probed rate."""
    tot = r["blocks"] * r["block_bytes"]
    gbs = r["probe"]["h2d_plus_d2h_concurrent_gbs_each_direction_all_ranks"]
    return {"host_link_probe": r["probe"], "floor_ms_per_batch": round(tot / (gbs * 1e9) * 1e3, 2),
            "note": "no kernel time included: the batch cannot finish faster than its bytes cross the host links"}


def c4_subrecord(r, world):
    tot = r["blocks"] * r["block_bytes"]
    return {"workload": WORKLOAD_C4, "n_gpus": world, "value": round(tot / 1e6 / (r["dev_ms"] / 1e3), 1), "unit": "MB/s",
            "ms_per_batch": round(r["dev_ms"], 2),
            "e2e": {"value": round(tot / 1e6 / (r["e2e_ms"] / 1e3), 1), "unit": "MB/s", "ms_per_batch": round(r["e2e_ms"], 2),
                    "api": "libsais_cuda_bwt_batch(pinned host blocks)", "h2d_bytes_per_batch": tot, "d2h_bytes_per_batch": tot,
                    "host_threads_per_gpu": r["lanes"], "lanes_tried_ms": r["lanes_tried_ms"], "host_link_floor": e2e_floor(r)},
            "every_block_verified": r["parity_ok"], "rounds": r["rounds"]}


# --------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--n", type=int, default=N_FULL, help="text bytes per step of the N = 1 workload (default: the 256 MiB config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipelined", action="store_true", help="(kept for compatibility; the two-thread measurement moved into the c4 record)")
    ap.add_argument("--no-profile", action="store_true", help="do not record per-kernel CUDA events in the timed region")
    ap.add_argument("--skip-c3", action="store_true")
    ap.add_argument("--skip-c4", action="store_true")
    ap.add_argument("--c4-blocks", type=int, default=C4_BLOCKS)
    ap.add_argument("--c4-block-bytes", type=int, default=C4_BLOCK_BYTES)
    ap.add_argument("--lanes", type=int, default=3, help="host threads / contexts per GPU of the batch entry point")
    ap.add_argument("--lanes-sweep", default="2,4", help="other lane counts tried by the end-to-end arm of the batch (best is reported)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import libsais_b200
    from libsais_b200 import gen

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or libsais_b200.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; libsais_cuda has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = numa_bind(local) if world > 1 else "single process"
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    peak, peak_src = peaks()
    ctx = libsais_b200.Context(local)
    sampler = ClockSampler(local) if rank == 0 else None

    if world > 1:
        # ------------------------------------------------------------------ configs[3]: strong scaling over the ranks
        r = run_c4(args, ctx, dev, local, rank, world, dist, barrier, sampler, reps=max(1, args.steps), warmup=max(args.warmup, 0))
        clocks = sampler.stop() if sampler else None
        if rank == 0:
            tot = r["blocks"] * r["block_bytes"]
            steps_dev = r["reps"] * r["blocks_this_rank"]
            roof, kern = roofline_record(r["agg"], steps_dev, r["dev_ms"] * r["reps"], peak, peak_src, KERNEL_NAMES)
            out = {"metric": "bwt_construction_throughput", "value": round(tot / 1e6 / (r["dev_ms"] / 1e3), 1), "unit": "MB/s", "n_gpus": world,
                   "steps": args.steps, "steps_timed": r["reps"], "warmup": args.warmup, "ms_per_step": round(r["dev_ms"], 3),
                   "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                   "config": {"workload": WORKLOAD_C4, "text_bytes": r["block_bytes"], "blocks": r["blocks"],
                              "step": "one pass over the whole 64-block batch (8 GiB of input); rank r processes blocks r, r+N, ...",
                              "parallelism": "one process per GPU, no collective on the data path; %d host threads/contexts per GPU in the end-to-end arm" % args.lanes,
                              "l2": "every block (128 MiB text, 1.5 GiB key/value stream) is larger than the 126 MB L2; no flush needed",
                              "numa": numa, "rounds_last_block": r["rounds"],
                              "note": "the N = 1 line of this script reports configs[1] (one 256 MiB text) as its headline and this batch on one GPU in its `c4` record"},
                   "e2e": {"value": round(tot / 1e6 / (r["e2e_ms"] / 1e3), 1), "unit": "MB/s", "h2d_bytes_per_step": tot, "d2h_bytes_per_step": tot,
                           "ms_per_step": round(r["e2e_ms"], 3), "api": "libsais_cuda_bwt_batch(pinned host blocks), one call per rank",
                           "host_threads_per_gpu": r["lanes"], "lanes_tried_ms": r["lanes_tried_ms"], "host_link_floor": e2e_floor(r)},
                   "gpu_launches": int(r["launches"]) * world, "every_block_verified": r["parity_ok"],
                   "roofline": roof, "kernels": kern, "clocks": clocks}
            print(json.dumps(out))
        dist.barrier()
        dist.destroy_process_group()
        return

    # ---------------------------------------------------------------------- configs[1]: one 256 MiB text
    n = args.n
    T = gen.rand_bytes(SEED, n)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    dT = torch.from_numpy(T).to(dev)
    dU = torch.empty(n, dtype=torch.uint8, device=dev)
    hT = torch.from_numpy(T).pin_memory()
    hU = torch.empty(n, dtype=torch.uint8).pin_memory()
    hA = np.empty(1, dtype=np.int32)                          # required non-NULL, never touched
    torch.cuda.synchronize()

    # ---- device-resident arm
    ctx.set_profiling(not args.no_profile)
    primary = None
    for _ in range(args.warmup):
        primary = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), n)
        assert primary > 0, "bwt_dev failed: %d (cuda error %d)" % (primary, ctx.last_error())
    barrier()
    if sampler:
        sampler.mark_start()
    agg, launches = {}, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        rc = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), n)
        assert rc == primary
        launches += accumulate(agg, ctx.stats())
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    if sampler:
        sampler.mark_stop()
    clocks = sampler.stop() if sampler else None
    rounds = ctx.stats()["rounds"]

    # ---- end-to-end arm: the drop-in call with host buffers
    ctx.set_profiling(False)
    for _ in range(2):
        rc = ctx.bwt_ptr(hT.data_ptr(), hU.data_ptr(), hA.ctypes.data, n)
        assert rc == primary
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = ctx.bwt_ptr(hT.data_ptr(), hU.data_ptr(), hA.ctypes.data, n)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    assert rc == primary
    # the two arms agree, and the result inverts back to the input (size-independent property)
    assert torch.equal(hU.to(dev), dU), "host-API BWT differs from device-API BWT"
    dBack = torch.empty(n, dtype=torch.uint8, device=dev)
    assert ctx.unbwt_dev(dU.data_ptr(), dBack.data_ptr(), n, primary) == 0 and torch.equal(dBack, dT), "unbwt(bwt(T)) != T"
    gpu_U = hU.numpy().copy()
    del dBack, dT, dU, hT, hU
    torch.cuda.empty_cache()

    value = n * args.steps / 1e6 / (dev_ms / 1e3)
    e2e = n * args.steps / 1e6 / (e2e_ms / 1e3)
    roof, kern = (roofline_record(agg, args.steps, dev_ms, peak, peak_src, KERNEL_NAMES) if agg and not args.no_profile else (None, {}))
    out = {
        "metric": "bwt_construction_throughput", "value": round(value, 1), "unit": "MB/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD_C2 if n == N_FULL else "libsais_bwt + primary index, %d MiB iid random bytes per step" % (n >> 20),
                   "text_bytes": n, "parallelism": "1 GPU",
                   "l2": "inputs (%d MiB text, %d MiB key/value stream) larger than the 126 MB L2; no flush needed" % (n >> 20, (n * 12) >> 20),
                   "rounds": rounds, "timed_region": "per-kernel CUDA events recorded inside it (profiling on)" if not args.no_profile else "profiling off"},
        "e2e": {"value": round(e2e, 1), "unit": "MB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": n + 4,
                "ms_per_step": round(e2e_ms / args.steps, 3), "api": "libsais_bwt_ctx(host pinned T, U), one host thread",
                "note": "H2D of the whole text first (every sort pass needs the histogram of the whole text); the D2H overlaps the sort: the "
                        "output buffer is pinned, so the rows of settled slots leave chunk by chunk while the remaining buckets are sorted, and "
                        "the bucket of suffix 0 and the few rows settled after round 0 follow (sa_core.cu streamed rows; api.cu bwt_body)"},
        "gpu_launches": int(launches), "roofline": roof, "kernels": kern, "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_c2(T, gpu_U, primary)
    del T, gpu_U
    if not args.skip_c4:
        try:
            r = run_c4(args, ctx, dev, local, 0, 1, None, barrier, None, reps=1)
            out["c4"] = c4_subrecord(r, 1)
        except Exception as e:                                # a sub-record must never take the headline down
            out["c4"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if not args.skip_c3:
        try:
            out["c3"] = run_c3(ctx, dev, peak)
        except Exception as e:
            out["c3"] = {"error": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
