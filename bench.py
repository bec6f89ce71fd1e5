#!/usr/bin/env python
"""bench.py -- BMU searches/s of the batch winner search on synthetic data (BASELINE.json
configs[2]: 10 M vectors x 64-dim vs a 100x100 map, qerror + visual style consumers).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's C code on host cores

One "step" = one pass of the hot path over the batch: the 10 M rows are cut into contiguous
shards, one per GPU (STRONG scaling: the total is fixed); every rank searches its shard against
the replicated codebook (no data-path collective), reduces its qerror sum / found count / BMU
histogram on the device, and the small statistics vector is combined by ONE grouped NCCL
all-reduce issued by the library.  For N > 1 a weak-scaling figure (10 M rows on
every GPU) is reported beside it under "weak".

JSON keys follow the driver contract; `value` is device-resident (inputs already in HBM),
`e2e` goes through bmu_search() with plain pageable HOST buffers (H2D + D2H inside the timed
region; the library stages them through its pinned ring).
The oracle/ directory is only used for the cpu_baseline leg and for --impl reference.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: rows IN TOTAL (sharded over the GPUs), D, M, xdim, k
    "c3": dict(rows=10_000_000, D=64, M=10_000, xdim=100, k=1,
               desc="synthetic batch winner search: 10M x 64-dim vs 100x100 map (BASELINE.json configs[2])"),
    "c4": dict(rows=1_000_000, D=512, M=4096, xdim=64, k=1,
               desc="synthetic high-dim: 1M x 512-dim vs 4096-unit codebook (BASELINE.json configs[3])"),
}

MASK64 = (1 << 64) - 1

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the
# committed ncu --set full capture of exactly this shape (profiles/): (kernel, rows, D, M) -> bytes
NCU_TRAFFIC = {("k2_rec_kernel", 10_000_000, 64, 10_000): 4423282000 + 122659584}   # profiles/r02_k2_rec_ncu_full.txt
# (fp16 A image 1.6 GB + the FP32 rows 2.56 GB the fused re-rank reads + codebook misses)


# ------------------------------------------------------------------ synthetic data
def _s64(v):
    v &= MASK64
    return v - (1 << 64) if v >= (1 << 63) else v


def synth_torch(seed, start, count, device):
    """counter-based generator: splitmix64(seed, index) -> 24-bit uniform in [0,1)"""
    import torch
    z = torch.arange(start, start + count, dtype=torch.int64, device=device)
    z = (z + _s64(seed * 0x632BE59BD9B4E019)) * _s64(0x9E3779B97F4A7C15)
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _s64(0x94D049BB133111EB)
    z = z ^ ((z >> 31) & ((1 << 33) - 1))
    return ((z >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / (1 << 24))


def synth_numpy(seed, start, count):
    with np.errstate(over="ignore"):
        z = np.arange(start, start + count, dtype=np.uint64)
        z = (z + np.uint64((seed * 0x632BE59BD9B4E019) & MASK64)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)) & np.uint64(0xFFFFFF)).astype(np.float32) * np.float32(1.0 / (1 << 24))


def synth_rows_torch(seed, row0, rows, D, device):
    import torch
    out = torch.empty((rows, D), dtype=torch.float32, device=device)
    step = 1 << 20
    for r in range(0, rows, step):
        n = min(step, rows - r)
        out[r:r + n] = synth_torch(seed, (row0 + r) * D, n * D, device).view(n, D)
    return out


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (the timed
    region of the search is tens of milliseconds, shorter than one nvidia-smi call); nvidia-smi -lms as
    the fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu):
        self.sm, self.mx, self.mask, self.p, self.thread = [], None, 0, None, None
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def _poll(self):
        nv = self.nv
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                if reasons:
                    self.mask |= int(reasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out["sm_mhz"] = float(np.median(self.sm))
                out["sm_max_mhz"] = self.mx
                out["samples"] = len(self.sm)
                out["source"] = "nvml"
            out["reasons"] = sorted(n for n, b in self.BITS.items() if self.mask & b)
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 6:
                continue
            try:
                sm.append(float(t[0]))
                mx.append(float(t[1]))
            except ValueError:
                continue
            for nme, v in zip(names, t[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.f.name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
            out["source"] = "nvidia-smi"
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------ CPU legs (oracle/ is allowed here only)
def _cpu_worker(args):
    kind, codes, data, k = args
    from oracle.pyoracle import Oracle, Reference
    if kind == "reference":
        t, _ = Reference().search_time_only(codes, data, k)
        return t
    o = Oracle()
    t0 = time.perf_counter()
    o.search(codes, data, k)
    return time.perf_counter() - t0


def cpu_searches_per_s(w, rows_per_core, cores, k):
    """time the reference's own find_winner_euc/knn (oracle/_ref when it was built, else the
    oracle port) on `cores` forked processes over disjoint slices of the same workload"""
    import multiprocessing as mp
    from oracle.pyoracle import Reference, build
    kind = "reference" if Reference.available() else "port"
    if kind == "port":
        build(ref=False)
    codes = synth_numpy(2, 0, w["M"] * w["D"]).reshape(w["M"], w["D"])
    jobs = []
    for c in range(cores):
        data = synth_numpy(1, c * rows_per_core * w["D"], rows_per_core * w["D"]).reshape(rows_per_core, w["D"])
        jobs.append((kind, codes, data, k))
    t0 = time.perf_counter()
    if cores == 1:
        times = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            times = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    # all slices run concurrently: throughput = total rows / slowest slice
    return rows_per_core * cores / max(times), kind, wall


def main_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # ~1.6 k searches/s/core at C3: bound one step to a few seconds
    rows_per_core = args.ref_rows or max(64, int(4.0 * 1600 * (10_000 * 64) / (w["M"] * w["D"])))
    vals = []
    for it in range(args.warmup + args.steps):
        v, kind, _ = cpu_searches_per_s(w, rows_per_core, cores, w["k"])
        if it >= args.warmup:
            vals.append(v)
    value = float(len(vals) / sum(1.0 / v for v in vals))
    line = {
        "impl": "reference", "metric": "BMU searches/s", "value": value, "unit": "searches/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * rows_per_core * cores / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "rows_per_step": rows_per_core * cores, "D": w["D"],
                   "M": w["M"], "k": w["k"]},
        "cpu_baseline": {"value": value, "unit": "searches/s", "cores": cores, "kind": kind,
                         "sample": "%d rows per core x %d forked processes of the same synthetic workload per step"
                                   % (rows_per_core, cores)},
        "e2e": {"value": value, "unit": "searches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    sys.stdout.flush()
    return 0


# ------------------------------------------------------------------ GPU arm
def main_gpu(args, w):
    import torch
    import torch.distributed as dist
    import som_lvq_pak_b200 as bmu
    from som_lvq_pak_b200 import _lib
    from som_lvq_pak_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the BMU engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bmu.init(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        D.comm_init()                 # the library's own NCCL communicator (id moved by torch.distributed)
    lib = _lib.load()
    _lib.check(lib.bmu_set_search_path({"auto": 0, "exact": 1, "filter": 2}[args.path]))

    # STRONG scaling (the named config): `total` rows in all, rank r searches the contiguous shard
    # [lo, hi) the library's own rule gives it (bmu_multi_shard_bounds)
    total, D_, M, k = args.rows or w["rows"], w["D"], w["M"], w["k"]
    Dm = D_
    lo, hi = D.shard_bounds(total, rank, world)
    rows = hi - lo
    stream = torch.cuda.current_stream().cuda_stream
    # codebook: generated on rank 0, replicated by ONE ncclBroadcast inside the library
    codes = synth_rows_torch(2, 0, M, Dm, dev) if rank == 0 else torch.empty((M, Dm), device=dev)
    _lib.check(lib.bmu_comm_broadcast_dev(codes.data_ptr(), M * Dm * 4, 0, stream))
    torch.cuda.synchronize()
    cb = lib.bmu_codebook_create_dev(codes.data_ptr(), M, Dm)
    if not cb:
        raise SystemExit("bmu_codebook_create_dev: " + lib.bmu_last_error().decode())
    data = synth_rows_torch(1, lo, rows, Dm, dev)
    ss = D.ShardedSearch(cb, M, rows, k, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)          # max over ranks
        return float(t[0])

    warm = max(args.warmup, 3)
    for _ in range(warm):
        ss.step(data.data_ptr())
    barrier()
    kms = (ctypes.c_float * 8)()
    launches0 = lib.bmu_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(lambda: ss.step(data.data_ptr()), args.steps)
    launches = lib.bmu_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    # duration of the dominant kernel, measured live with CUDA events on the launching stream: events
    # bracket every kernel inside the library and are kept for the last 32 calls, so the figures below
    # are the AVERAGE over the timed steps (the board's power cap makes late steps slower than early ones)
    nhist = min(args.steps, 32)
    kernel_ms = [0.0] * 8
    for back in range(nhist):
        lib.bmu_search_kernel_ms_history(back, kms)
        for i in range(8):
            kernel_ms[i] += float(kms[i]) / nhist
    t = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    launches = int(t[0])
    value = total * args.steps / (ms * 1e-3)
    qsum, nfound, hist = ss.totals()
    bd = bmu.last_search_breakdown()
    idx, diff = ss.idx, ss.diff

    # ---- e2e: the reference-facing C-ABI call with plain HOST buffers (pageable numpy arrays, what a C
    # host's calloc gives, datafile.c:472), host<->device copies inside the timed region
    h_data = np.empty((rows, Dm), np.float32)
    step_rows = 1 << 20
    for r in range(0, rows, step_rows):
        h_data[r:r + step_rows] = data[r:r + step_rows].cpu().numpy()
    h_idx = np.empty((rows, k), np.int32)
    h_diff = np.empty((rows, k), np.float32)
    h_nf = np.empty(rows, np.int32)

    def e2e_step():
        _lib.check(lib.bmu_search(cb, h_data.ctypes.data, None, rows, k, h_idx.ctypes.data,
                                  h_diff.ctypes.data, h_nf.ctypes.data))

    def wall(fn, steps):
        for _ in range(max(args.warmup, 3)):                   # untimed warm-up calls, like the device-resident leg (the
            fn()                                               # first ones allocate the pinned ring and fault in the
        barrier()                                              # caller's result arrays)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0])

    e2e_steps = max(1, min(args.steps, 3))
    e2e_s = wall(e2e_step, e2e_steps)
    e2e_value = total * e2e_steps / e2e_s
    same = bool((torch.from_numpy(h_idx).to(dev) == idx).all()) and bool((torch.from_numpy(h_diff).to(dev) == diff).all())
    # the same call on page-locked buffers (bmu_host_register): the DMA reads the caller's arrays directly
    e2e_pinned = None
    if not args.no_pinned:
        regs = [(h_data.ctypes.data, h_data.nbytes), (h_idx.ctypes.data, h_idx.nbytes),
                (h_diff.ctypes.data, h_diff.nbytes), (h_nf.ctypes.data, h_nf.nbytes)]
        ok = all(lib.bmu_host_register(p_, n_) == 0 for p_, n_ in regs)
        if ok:
            e2e_pinned = total * e2e_steps / wall(e2e_step, e2e_steps)
        for p_, _ in regs:
            lib.bmu_host_unregister(p_)
    del h_data

    # ---- weak scaling beside the strong headline (N > 1): `total` rows on EVERY GPU
    weak = None
    if world > 1 and not args.no_weak:
        del data, ss
        torch.cuda.empty_cache()
        wdata = synth_rows_torch(1, rank * total, total, Dm, dev)
        ws = D.ShardedSearch(cb, M, total, k, dev)
        for _ in range(2):
            ws.step(wdata.data_ptr())
        wsteps = max(1, min(args.steps, 3))
        wms = timed(lambda: ws.step(wdata.data_ptr()), wsteps)
        weak = {"value": world * total * wsteps / (wms * 1e-3), "unit": "searches/s", "ms_per_step": wms / wsteps,
                "rows_per_gpu": total, "steps": wsteps}
        del wdata, ws

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        info = bmu.device_info()
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        fp32_peak = info["sm_count"] * 128 * sm_max * 1e6 / 1e12          # T lane-ops/s, nominal
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        used_k2 = bd["k2_certified"] + bd["k2_failed"] > 0
        names = ("k1_data_prep", "k1_fast", "k1_warp", "k1_seq", "k2_row_prep", "k2_gemm", "k2_rerank", "k2_lists")
        step_ms = {n: v for n, v in zip(names, kernel_ms) if (n.startswith("k2") == used_k2)}
        hbm_bytes = rows * (4.0 * Dm + 12.0 * k)
        if used_k2:
            kp = ((Dm + 7) // 8 * 8 + 3 + 15) // 16 * 16                   # fp16 operand K: D + 3 norm columns
            k_ms = kernel_ms[5]
            kname = "k2_rec_kernel" if (k == 1 and kp <= 96) else "k2_gemm_kernel"
            flop = 2.0 * M * Dm * rows                                     # 2*M*D per search
            achieved = flop / (k_ms * 1e-3) / 1e12
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            m_pad = (M + 255) // 256 * 256
            issued = achieved * (kp / Dm) * (m_pad / M)
            traffic = NCU_TRAFFIC.get((kname, rows, Dm, M))
            roof = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "kernel_ms": k_ms, "step_kernels_ms": step_ms,
                    "note": "algorithmic 2*M*D flop per search over the live CUDA-event duration of %s (rank 0's shard) "
                            "vs the sustained fp16/bf16 cuBLAS peak (%s); the kernel issues %.2fx that many MMA flops "
                            "(K %d -> %d: 3 norm columns + padding to 16; M %d -> %d), so the tensor pipe itself "
                            "runs at %.3f of that peak.  For k = 1 the timed kernel also contains the exact re-rank "
                            "(4 extra warps; FMA-pipe work that replaced a separate 2.3 ms kernel), which the flop "
                            "count above does not credit.  traffic = ncu dram bytes of one launch (profiles/), null "
                            "when no capture exists for this shape"
                            % (kname, "measured" if "bf16_tflops_sustained" in peaks else "fallback",
                               (kp / Dm) * (m_pad / M), Dm, kp, M, m_pad, issued / peak),
                    "mma_issued_tflops": issued,
                    "rows_certified": bd["k2_certified"], "rows_redone_exactly": bd["k2_failed"]}
            search_path = "filter (K2 tcgen05 GEMM + exact re-rank, K1 for uncertified rows)"
        else:
            k_ms = kernel_ms[1] if kernel_ms[1] > 0 else kernel_ms[2]
            kname = "k1_fast_kernel" if kernel_ms[1] > 0 else "k1_warp_kernel"
            flop = 3.0 * M * Dm * rows                                     # 3*M*D per search
            achieved = flop / (k_ms * 1e-3) / 1e12
            roof = {"bound": "fp32_issue", "kernel": kname, "achieved": achieved, "peak": fp32_peak,
                    "unit": "TFLOP/s", "frac": achieved / fp32_peak, "traffic": None, "kernel_ms": k_ms,
                    "step_kernels_ms": step_ms,
                    "note": "exact path is bounded by the non-FMA FP32 issue rate, not HBM (SURVEY.md 8d): 3*M*D "
                            "lane-ops per search; peak = SMs*128*sm_max_clock (nominal; tools/ubench/fp32_issue "
                            "measured 36.8 T lane-ops/s on this pool)"}
            search_path = "exact (K1)"
        roof["hbm"] = {"achieved": hbm_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": hbm_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                       "peak_source": "measured" if "hbm_gbs" in peaks else "fallback"}
        line = {
            "metric": "BMU searches/s", "value": value, "unit": "searches/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": w["desc"], "rows_total": total, "rows_per_gpu": rows, "D": Dm, "M": M, "k": k,
                       "parallelism": "rows sharded x%d inside libbmu_b200 (contiguous slices), codebook replicated by "
                                      "one ncclBroadcast, ONE grouped ncclAllReduce per step of 1 double + %d int64 "
                                      "(sum sqrt(diff); n_found, BMU histogram)" % (world, 1 + M),
                       "l2": "inputs (%.2f GB per GPU) are larger than L2 (126 MB)" % (rows * Dm * 4 / 1e9),
                       "search_path": search_path},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "searches/s", "h2d_bytes_per_step": total * Dm * 4,
                    "d2h_bytes_per_step": total * k * 8 + total * 4, "steps": e2e_steps,
                    "host_buffers": "pageable (numpy), staged through the library's pinned ring by its copy threads",
                    "value_page_locked_buffers": e2e_pinned,
                    "matches_device_resident_run": same},
            "gpu_launches": launches,
            "roofline": roof,
            "result_check": {"mean_qerror": qsum / max(nfound, 1), "n_found": nfound,
                             "hist_total": int(hist.sum())},
        }
        if weak:
            line["weak"] = weak
        if world == 1 and not args.no_vsom:
            line["vsom"] = vsom_c5(bmu, args)
        if world == 1 and not args.no_c4 and args.workload == "c3":
            try:
                del data, ss
            except NameError:
                pass
            del idx, diff
            torch.cuda.empty_cache()
            line["c4"] = c4_extra(bmu, lib, _lib, dev, peaks)
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            rpc = args.ref_rows or max(64, int(10.0 * 1600 * (10_000 * 64) / (M * Dm)))
            v, kind, wall_s = cpu_searches_per_s(w, rpc, cores, k)
            line["cpu_baseline"] = {"value": v, "unit": "searches/s", "cores": cores, "kind": kind,
                                    "sample": "first %d rows per core x %d forked processes, same codebook "
                                              "(reference find_winner_euc, gcc -O3), %.1f s wall"
                                              % (rpc, cores, wall_s)}
        print(json.dumps(line))
        sys.stdout.flush()
    lib.bmu_codebook_destroy(cb)
    if world > 1:
        lib.bmu_comm_destroy()
        dist.destroy_process_group()
    return 0


def c4_extra(bmu, lib, _lib, dev, peaks):
    """BASELINE.json configs[3]: 1 M x 512-dim vs a 4096-unit codebook, k = 1 (accuracy / classify)
    and k = 5 (knntest) -- the long-K streaming tcgen05 kernel.  Device-resident, CUDA events."""
    import torch
    rows, D, M = 1_000_000, 512, 4096
    codes = synth_rows_torch(2, 0, M, D, dev)
    data = synth_rows_torch(3, 0, rows, D, dev)
    cb = lib.bmu_codebook_create_dev(codes.data_ptr(), M, D)
    out = {"config": "synthetic high-dim: 1M x 512-dim vs 4096-unit codebook (BASELINE.json configs[3])"}
    kms = (ctypes.c_float * 8)()
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    for k in (1, 5):
        idx = torch.empty((rows, k), dtype=torch.int32, device=dev)
        diff = torch.empty((rows, k), dtype=torch.float32, device=dev)
        nf = torch.empty(rows, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            _lib.check(lib.bmu_search_dev(cb, data.data_ptr(), None, rows, k, idx.data_ptr(), diff.data_ptr(),
                                          nf.data_ptr(), stream))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps = 5
        for _ in range(steps):
            _lib.check(lib.bmu_search_dev(cb, data.data_ptr(), None, rows, k, idx.data_ptr(), diff.data_ptr(),
                                          nf.data_ptr(), stream))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        avg = [0.0] * 8
        for back in range(steps):                                          # average over the timed steps
            lib.bmu_search_kernel_ms_history(back, kms)
            for i in range(8):
                avg[i] += float(kms[i]) / steps
        bd = bmu.last_search_breakdown()
        gemm_ms = avg[5]
        out["k%d" % k] = {"value": rows / (ms * 1e-3), "unit": "searches/s", "ms_per_step": ms,
                          "kernel_ms": {"k2_row_prep": avg[4], "k2_gemm": gemm_ms, "k2_rerank": avg[6],
                                        "k2_lists": avg[7]},
                          "gemm_tflops_algorithmic": 2.0 * M * D * rows / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None,
                          "gemm_frac_of_sustained_peak": (2.0 * M * D * rows / (gemm_ms * 1e-3) / 1e12 / peak) if gemm_ms > 0 else None,
                          "rows_certified": bd["k2_certified"], "rows_redone_exactly": bd["k2_failed"]}
        del idx, diff, nf
    lib.bmu_codebook_destroy(cb)
    return out


def vsom_c5(bmu, args):
    """second headline metric: vsom steps/s on BASELINE.json configs[4] (256x256 hexa gaussian map,
    128-dim, rlen 1e6; SURVEY 8d parameters).  A bounded prefix of the 1e6-step schedule is run
    on the GPU (CUDA events inside the library); the CPU leg runs the reference's own
    som_training (oracle/_ref) on one core for a few steps (training is sequential)."""
    N, D, xdim, ydim, length = 100_000, 128, 256, 256, 1_000_000
    data = synth_numpy(4, 0, N * D).reshape(N, D)
    codes = synth_numpy(5, 0, xdim * ydim * D).reshape(xdim * ydim, D)
    order = bmu.rand_order(N, 3)
    steps = args.vsom_steps
    s, ta, tr = bmu.som_schedule(0, steps, length, 0.05, 100.0, bmu.ALPHA_LINEAR, N, order)
    t = bmu.Trainer(codes, data)
    t.set_som(xdim, ydim, bmu.TOPOL_HEXA, bmu.NEIGH_GAUSSIAN)
    t.steps(s[:64], ta[:64], tr[:64])              # warm-up launch
    t.close()
    t = bmu.Trainer(codes, data)
    t.set_som(xdim, ydim, bmu.TOPOL_HEXA, bmu.NEIGH_GAUSSIAN)
    t0 = time.perf_counter()
    t.steps(s, ta, tr)
    wall = time.perf_counter() - t0
    ms = t.last_ms()
    t.close()
    M = xdim * ydim
    info = bmu.device_info()
    try:
        sm_max = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("sm_max_mhz", 1965.0)
    except Exception:
        sm_max = 1965.0
    fp32_peak = info["sm_count"] * 128 * sm_max * 1e6 / 1e12
    ach = 6.0 * M * D * steps / (ms * 1e-3) / 1e12
    out = {"metric": "vsom steps/s", "value": steps / (ms * 1e-3), "unit": "steps/s",
           "e2e_value": steps / wall, "steps": steps, "us_per_step": 1e3 * ms / steps,
           "config": "256x256 hexa gaussian map, 128-dim, rlen 1e6 schedule (%s), alpha 0.05 "
                     "linear, radius 100, -rand 3 order, 100000 x 128 synthetic data"
                     % ("all 1e6 steps" if steps == length else "first %d steps" % steps),
           "lane_ops_per_step": 6.0 * M * D,
           "roofline": {"bound": "fp32_issue", "kernel": "k3_som_fused_kernel", "achieved": ach, "peak": fp32_peak,
                        "unit": "T lane-ops/s", "frac": ach / fp32_peak,
                        "note": "6*M*D non-FMA FP32 lane-ops per gaussian step (search 3*M*D + update 3*M*D, SURVEY 8d) "
                                "over the CUDA-event duration of the one persistent launch, vs SMs*128*sm_max_clock; the "
                                "step is a latency chain (grid-wide winner exchange -> lattice weights -> one pass), so "
                                "this fraction is what the chain leaves, not a pipe limit"}}
    if not args.no_cpu:
        from oracle.pyoracle import Reference, Oracle
        nref = 30
        t0 = time.perf_counter()
        if Reference.available():
            Reference().som_train(codes, data, xdim, ydim, 3, 2, nref, 0.05, 100.0, 1, rand_seed=3)
            kind = "reference"
        else:
            Oracle().som_train(codes, data, xdim, ydim, 3, 2, nref, 0.05, 100.0, 1, order=order)
            kind = "port"
        dt = time.perf_counter() - t0
        # list building / copies are outside the loop of interest: subtract a 1-step run
        t0 = time.perf_counter()
        if kind == "reference":
            Reference().som_train(codes, data, xdim, ydim, 3, 2, 1, 0.05, 100.0, 1, rand_seed=3)
        else:
            Oracle().som_train(codes, data, xdim, ydim, 3, 2, 1, 0.05, 100.0, 1, order=order)
        dt1 = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": (nref - 1) / max(dt - dt1, 1e-9), "unit": "steps/s", "cores": 1,
                               "kind": kind, "sample": "%d steps of som_training (rlen %d) minus a 1-step run"
                                                       % (nref, nref)}
    return out


def main():
    # exactly ONE line on stdout: libraries print there too (NCCL's version banner), so file descriptor 1 is
    # pointed at stderr for the run and the JSON line goes to the saved descriptor
    real_stdout = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override the total number of rows (debug)")
    ap.add_argument("--ref-rows", type=int, default=0, help="override CPU sample rows per core")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-vsom", action="store_true", help="skip the vsom (configs[4]) extra metric")
    ap.add_argument("--no-c4", action="store_true", help="skip the high-dim (configs[3]) extra metric")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling extra (N > 1)")
    ap.add_argument("--no-pinned", action="store_true", help="skip the page-locked variant of the e2e leg")
    ap.add_argument("--vsom-steps", type=int, default=1_000_000)
    ap.add_argument("--path", default="auto", choices=["auto", "exact", "filter"],
                    help="search kernels: auto (default), exact = K1 only, filter = K2 forced")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return main_reference(args, w)
    return main_gpu(args, w)


if __name__ == "__main__":
    sys.exit(main())
