#!/usr/bin/env python
"""bench.py -- headline benchmark of libsais_cuda (contract: see the task statement / DESIGN.md §6).

Workload (BASELINE.json configs[1]): libsais_bwt + primary index of a 256 MiB synthetic
random-byte text (splitmix64 generator, seed 2 + rank), one text per GPU per step.
  value : MB/s of input, device-timed with CUDA events on the context's stream, input and
          output resident in HBM (libsais_cuda_bwt_dev).
  e2e   : same metric through the reference-facing C-ABI call libsais_bwt_ctx() with HOST
          (pinned) buffers: H2D of the text and D2H of the BWT inside the timed region.
  roofline : the dominant kernel (onesweep digit pass), algorithmic bytes / CUDA-event time
          measured live inside the timed region, against MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline : the unmodified reference (oracle/_ref, libsais_bwt_omp) on this box's host
          cores, on a bounded sample of the same generator.
N > 1 (torchrun): independent texts, one per rank, no collective on the data path -> "weak".
`--impl reference` times the reference's own CPU implementation instead (rank 0 only).
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
CPU_SAMPLE = 1 << 26        # 64 MiB sample of the same generator for the CPU arm
SEED = 2
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


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


def ref_lib():
    p = os.path.join(ROOT, "oracle", "_ref", "libsais_ref.so")
    if not os.path.exists(p):
        if os.path.isdir("/root/reference/src"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
        else:
            return None
    lib = C.CDLL(p)
    lib.libsais_bwt_omp.restype = C.c_int32
    return lib


def cpu_bwt_mbs(lib, T, threads, reps=1):
    """Reference libsais_bwt_omp wall-clock MB/s (best of reps)."""
    n = len(T)
    U = np.empty(n, dtype=np.uint8)
    A = np.empty(n, dtype=np.int32)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = lib.libsais_bwt_omp(T.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p),
                                 C.c_int32(n), C.c_int32(0), None, C.c_int32(threads))
        dt = time.perf_counter() - t0
        assert rc > 0
        best = dt if best is None else min(best, dt)
    return n / 1e6 / best, rc, U


def cpu_baseline(sample_n=CPU_SAMPLE):
    from libsais_b200 import gen
    lib = ref_lib()
    if lib is None:
        return None
    T = gen.rand_bytes(SEED, sample_n)
    ncpu = os.cpu_count() or 1
    best = (0.0, 1)
    for th in sorted({min(8, ncpu), min(16, ncpu), min(32, ncpu), ncpu}):
        v, _, _ = cpu_bwt_mbs(lib, T, th)
        if v > best[0]:
            best = (v, th)
    return {"value": round(best[0], 2), "unit": "MB/s", "cores": best[1], "kind": "reference",
            "sample": "libsais_bwt_omp (libsais 2.10.4, gcc -O3 -march=x86-64-v3 -fopenmp) on the first %d MiB of the same "
                      "random-byte generator, best of threads in {8,16,32,%d}, host has %d logical CPUs" % (sample_n >> 20, ncpu, ncpu)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from libsais_b200 import gen
    lib = ref_lib()
    ncpu = os.cpu_count() or 1
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsais_ref.so missing and /root/reference absent"}))
        return
    n = CPU_SAMPLE
    T = gen.rand_bytes(SEED, n)
    threads = ncpu
    for _ in range(args.warmup):
        cpu_bwt_mbs(lib, T[: n // 8], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_bwt_mbs(lib, T, threads)
    dt = time.perf_counter() - t0
    v = n * args.steps / 1e6 / dt
    sample = "each step = libsais_bwt_omp on a %d MiB sample of the 256 MiB random-byte text, %d threads" % (n >> 20, threads)
    print(json.dumps({
        "impl": "reference", "metric": "bwt_construction_throughput", "value": round(v, 2), "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "libsais_bwt + primary index, 256 MiB iid random bytes (configs[1]); CPU arm runs a bounded sample", "sample_bytes": n},
        "cpu_baseline": {"value": round(v, 2), "unit": "MB/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": round(v, 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--n", type=int, default=N_FULL, help="text bytes per GPU per step (default: the 256 MiB config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipelined", action="store_true", help="skip the supplementary two-thread end-to-end measurement")
    ap.add_argument("--no-profile", action="store_true", help="do not record per-kernel CUDA events in the timed region")
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available() or libsais_b200.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; libsais_cuda has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    n = args.n
    T = gen.rand_bytes(SEED + rank, n)                       # one independent text per rank
    ctx = libsais_b200.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    dT = torch.from_numpy(T).to(dev)
    dU = torch.empty(n, dtype=torch.uint8, device=dev)
    hT = torch.from_numpy(T).pin_memory()
    hU = torch.empty(n, dtype=torch.uint8).pin_memory()
    hA = np.empty(1, dtype=np.int32)                          # required non-NULL, never touched
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm
    sampler = ClockSampler(local) if rank == 0 else None
    ctx.set_profiling(not args.no_profile)
    primary = None
    for _ in range(args.warmup):
        primary = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), n)
        assert primary > 0, "bwt_dev failed: %d (cuda error %d)" % (primary, ctx.last_error())
    barrier()
    if sampler:
        sampler.mark_start()
    agg = {}
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        rc = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), n)
        assert rc == primary
        st = ctx.stats()
        launches += st["total_launches"]
        for k, v in st["kernels"].items():
            a = agg.setdefault(k, {"launches": 0, "ms": 0.0, "bytes": 0.0})
            a["launches"] += v["launches"]; a["ms"] += v["ms"]; a["bytes"] += v["bytes"]
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
    e2e_s = time.perf_counter() - t0
    assert rc == primary
    # ---- supplementary: the reference's threading model (one context per host thread, include/libsais.h:53-55)
    # applied to ONE GPU: two host threads alternate texts, so one call's PCIe copies overlap the other's kernels
    piped_s = None
    if not args.no_pipelined:
        import threading
        ctx2 = libsais_b200.Context(local)
        hT2 = torch.from_numpy(T).pin_memory()
        hU2 = torch.empty(n, dtype=torch.uint8).pin_memory()
        per_thread = max(1, args.steps // 2)
        rcs = [0, 0]

        def worker(k, c, t_in, t_out):
            for _ in range(per_thread):
                rcs[k] = c.bwt_ptr(t_in.data_ptr(), t_out.data_ptr(), hA.ctypes.data, n)

        worker(1, ctx2, hT2, hU2)                       # warm the second context's workspace
        barrier()
        th = [threading.Thread(target=worker, args=(0, ctx, hT, hU)), threading.Thread(target=worker, args=(1, ctx2, hT2, hU2))]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        piped_s = time.perf_counter() - t0
        assert rcs[0] == primary and rcs[1] == primary and torch.equal(hU2, hU)
        ctx2.close()
    # the two arms agree, and the result inverts back to the input (size-independent property)
    assert torch.equal(hU.to(dev), dU), "host-API BWT differs from device-API BWT"
    dBack = torch.empty(n, dtype=torch.uint8, device=dev)
    assert ctx.unbwt_dev(dU.data_ptr(), dBack.data_ptr(), n, primary) == 0 and torch.equal(dBack, dT), "unbwt(bwt(T)) != T"

    times = torch.tensor([dev_ms, e2e_s * 1e3, (piped_s or 0.0) * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, piped_ms = float(times[0]), float(times[1]), float(times[2])

    if rank == 0:
        peak, peak_src = peaks()
        value = world * n * args.steps / 1e6 / (dev_ms / 1e3)
        e2e = world * n * args.steps / 1e6 / (e2e_ms / 1e3)
        dom = agg.get("sort_pass", {"launches": 0, "ms": 0.0, "bytes": 0.0})
        roof = None
        if dom["ms"] > 0:
            achieved = dom["bytes"] / (dom["ms"] / 1e3) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "sort_pass_traffic.json")
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get("dram_bytes_per_launch")
                except Exception:
                    pass
            roof = {"bound": "hbm", "kernel": "sort_pass_kernel<u64,u32> (onesweep digit pass)", "achieved": round(achieved, 1),
                    "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                    "algorithmic_bytes_per_launch": dom["bytes"] / max(dom["launches"], 1),
                    "avg_launch_ms": dom["ms"] / max(dom["launches"], 1), "launches": dom["launches"],
                    "note": "achieved = sum of algorithmic bytes / sum of CUDA-event time over ALL launches of the kernel in the timed "
                            "region (4 full-size passes of 2*n*12 B + 6 small passes of the ~64K unresolved suffixes per step); "
                            "traffic = dram read+write bytes of ONE full-size launch (ncu), to compare with full_size_algorithmic_bytes",
                    "full_size_algorithmic_bytes": 2 * n * 12,
                    "share_of_step": round(dom["ms"] / dev_ms, 4)}
        kern = {k: {"launches": v["launches"], "ms_per_step": round(v["ms"] / args.steps, 4),
                    "algo_gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in agg.items()}
        out = {
            "metric": "bwt_construction_throughput", "value": round(value, 1), "unit": "MB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "libsais_bwt + primary index, %d MiB iid random bytes per GPU per step (BASELINE configs[1]; splitmix64 seed 2+rank)" % (n >> 20),
                       "text_bytes": n, "parallelism": "independent texts, one per GPU, no collective" if world > 1 else "1 GPU",
                       "l2": "inputs (%d MiB text, %d MiB key/value stream) larger than the 126 MB L2; no flush needed" % (n >> 20, (n * 12) >> 20),
                       "rounds": rounds},
            "e2e": {"value": round(e2e, 1), "unit": "MB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": n + 4,
                    "ms_per_step": round(e2e_ms / args.steps, 3), "api": "libsais_bwt_ctx(host pinned T, U)"},
            "gpu_launches": int(launches), "roofline": roof, "kernels": kern, "clocks": clocks,
        }
        if piped_ms > 0:
            calls = 2 * max(1, args.steps // 2)
            out["e2e_pipelined"] = {"value": round(world * n * calls / 1e6 / (piped_ms / 1e3), 1), "unit": "MB/s", "calls": calls,
                                    "how": "two host threads per GPU, one context each (libsais's one-ctx-per-thread model), "
                                           "libsais_bwt_ctx on pinned host buffers; copies of one call overlap kernels of the other"}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
