#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for linfa_linalg_b200.

Step = one pass of the hot path over one synthetic input.  Headline workload (BASELINE.json
configs[1]): lower Cholesky of a 16384 x 16384 synthetic SPD f64 matrix on one B200, reported in
GFLOP/s (algorithmic n^3/3).  Cholesky of ONE matrix does not shard ("replicas only"):
with --gpus N every rank factors its own replica and `value` is the aggregate (weak scaling).
The sharded paths of the north star -- batched 32x32 f32 QR (batch-sharded, no collective) and
TSQR (row-sharded, one NCCL all-gather of the R factors) -- and the 16384^2 f64 QR are timed in the
same run and reported under "extras".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 16384] [--no-extras]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "QR/Cholesky f64 GFLOP/s vs FP64 peak @1/2/4/8 B200; batched QR matrices/s"
FP64_NOMINAL_TFLOPS = 37.0  # B200 FP64 (vector = tensor), nominal; MEASURED_PEAKS.json has no FP64 entry


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line).
    NVML (nvidia_ml_py) is polled from a thread every 10 ms so that even a 300 ms timed region gets
    tens of samples; `nvidia-smi -lms` needs ~1 s to start and is only the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []   # (sm_mhz, reasons bitmask)
        self.smax = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None
        self.proc = None
        self.lines = []

    def _phys_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except (ValueError, IndexError):
                pass
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._phys_index())
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self._phys_index())], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.01)

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1)
            nv = self.nvml
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                    "hw_power_brake": 0x80}
            sm = sorted(s for s, _ in self.samples)
            reasons = sorted({k for _, r in self.samples for k, b in bits.items() if r & b})
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 10 ms poll inside the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 50"}


# ---------------------------------------------------------------------------------------------
def workload_name(n):
    """config.workload, identical on both arms (the driver pairs the lines by it)."""
    return f"Cholesky {n}x{n} SPD f64, lower, in place (BASELINE configs[1])"


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the same metric/config, timed on the
    box's host cores.  The Rust crate cannot be built in this image (no rustc/cargo), so this runs
    the oracle port of src/cholesky.rs:51-83 -- single-threaded, like the reference."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import numpy as np
    import oracle as O
    n = args.ref_n
    g = np.random.default_rng(0x1F2E3D4C + 1).uniform(-1, 1, (n, n))
    s0 = (g + g.T) / 2 + n * np.eye(n)
    for _ in range(min(args.warmup, 1)):
        s = s0.copy(); O.cholesky(s)
    t = 0.0
    for _ in range(args.steps):
        s = s0.copy()
        t0 = time.perf_counter(); st, _ = O.cholesky(s); t += time.perf_counter() - t0
        assert st == 0
    flops = n ** 3 / 3.0
    val = flops * args.steps / t / 1e9
    # the rate of the unblocked row-Cholesky falls with n (the working set leaves the caches): show the trend so that the
    # reader can see which way the n = ref_n sample errs against the full-size workload (it flatters the CPU)
    trend = {}
    for tn in (1024, 2048, n):
        gg = np.random.default_rng(tn).uniform(-1, 1, (tn, tn))
        ss = (gg + gg.T) / 2 + tn * np.eye(tn)
        t0 = time.perf_counter(); O.cholesky(ss); dt = time.perf_counter() - t0
        trend[str(tn)] = round(tn ** 3 / 3.0 / dt / 1e9, 3)
    sample = (f"Cholesky f64 n={n} per step: a bounded SAMPLE of the n={args.n} workload (same algorithm, time scales as n^3; "
              f"the full size would take ~{t / args.steps * (args.n / n) ** 3 / 60:.0f} min per step on one core)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n), "parallelism": "host cores: 1 (the reference is single-threaded)", "sample": sample,
                   "same_config": False, "ref_n": n, "ratio_kind": "rate ratio (GFLOP/s over GFLOP/s), not a same-size time ratio",
                   "rate_trend_gflops_by_n": trend},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--ref-n", type=int, default=3072)      # ~3 s per step on one host core
    ap.add_argument("--cpu-n", type=int, default=5120)      # ~12 s of CPU work (the 10-30 s sample the contract asks for)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import linfa_linalg_b200 as L
    eng = L.Engine(local_rank)
    for kv in filter(None, os.environ.get("LFB_OPTS", "").split(",")):      # A/B runs: LFB_OPTS="chol_waves=1,gemm_tma2=0" python bench.py
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    lib = eng.lib
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    n = args.n
    seed = 0x1F2E3D4C + 1 + rank
    gen = torch.Generator(device=dev).manual_seed(seed)
    # S = (G + G^T)/2 + n I  (strictly diagonally dominant => SPD), built on the device
    S = torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen).mul_(2).sub_(1)
    S = S.add_(S.t().clone()).mul_(0.5)
    S.diagonal().add_(float(n))
    work = torch.empty_like(S)
    info = torch.zeros(1, dtype=torch.int64, device=dev)

    def chol_step():
        work.copy_(S)  # restore the input (in-place factorisation); 2 n^2 * 8 B of copy traffic, < 1 % of a step
        st = lib.lfb_cholesky_dev_f64(eng.h, C.c_void_p(work.data_ptr()), n, n, 0, C.c_void_p(info.data_ptr()))
        if st != 0:
            raise RuntimeError(f"lfb_cholesky_dev_f64 status {st}: {lib.lfb_last_error(eng.h)}")

    sampler = ClockSampler(local_rank)
    # warm-up first so that the clock samples cover the timed region only
    for _ in range(args.warmup):
        chol_step()
    barrier()
    launches0 = eng.launch_count
    sampler.start()
    ms = timed(chol_step, args.steps, 0)
    clocks = sampler.stop()
    launches = eng.launch_count - launches0  # kernels of liblinfa_b200 launched inside the timed region
    assert int(info.item()) == 0
    flops = n ** 3 / 3.0
    value = world * flops * args.steps / (ms * 1e-3) / 1e9

    # residual check of the last factor (size-independent property): ||S - L L^T||_F / ||S||_F on a block
    # (torch views the column-major buffer transposed: the lower factor L appears as triu(work) = L^T)
    Lf = torch.triu(work[:2048, :2048]).t()
    resid = float((Lf @ Lf.t() - S[:2048, :2048]).norm() / S[:2048, :2048].norm())

    # ---- roofline of the dominant kernel (FP64 GEMM launches), measured live with CUDA events ----
    peak_dmma = C.c_double(0.0); peak_dfma = C.c_double(0.0)
    lib.lfb_microbench_fp64(eng.h, 1, C.byref(peak_dmma))
    lib.lfb_microbench_fp64(eng.h, 0, C.byref(peak_dfma))
    lib.lfb_profile_begin(eng.h)
    chol_step()
    g_ms, g_fl, g_calls = C.c_double(0), C.c_double(0), C.c_int64(0)
    lib.lfb_profile_end(eng.h, C.byref(g_ms), C.byref(g_fl), C.byref(g_calls))
    peak_tf = max(peak_dmma.value, peak_dfma.value) / 1e3
    achieved_tf = g_fl.value / (g_ms.value * 1e-3) / 1e12 if g_ms.value > 0 else 0.0
    # DRAM traffic of the same launches from the committed ncu capture (bytes per launch, like `achieved`)
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "r2_chol16384_gemm_traffic.json")
    if n == 16384 and os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj["traffic_bytes_per_launch"]
        traffic_note = (f"ncu dram__bytes_read+write over the {tj['launches']} dgemm launches of one step: "
                        f"{(tj['dram_read_bytes'] + tj['dram_write_bytes']) / 1e9:.1f} GB = "
                        f"{(tj['dram_read_bytes'] + tj['dram_write_bytes']) / tj['algorithmic_flops'] * 1e3:.1f} B per kflop (compute bound)")
    roofline = {
        "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf if peak_tf > 0 else None, "traffic": traffic, "traffic_note": traffic_note,
        "kernel": "dgemm (FP64 DMMA.8x8x4) launches of one Cholesky step",
        "launches": int(g_calls.value), "kernel_ms_per_step": g_ms.value, "kernel_share_of_step": g_ms.value / (ms / args.steps),
        "peak_source": ("measured in this run by lfb_microbench_fp64 (register-resident DMMA / DFMA chains); "
                        "MEASURED_PEAKS.json has no FP64 entry"),
        "peak_dmma_tflops": peak_dmma.value / 1e3, "peak_dfma_tflops": peak_dfma.value / 1e3,
        "nominal_fp64_tflops": FP64_NOMINAL_TFLOPS,
        "step_tflops": flops / (ms / args.steps * 1e-3) / 1e12,
        "step_frac_of_measured_peak": flops / (ms / args.steps * 1e-3) / 1e12 / peak_tf if peak_tf > 0 else None,
        "step_frac_of_nominal_peak": flops / (ms / args.steps * 1e-3) / 1e12 / FP64_NOMINAL_TFLOPS,
    }

    # ---- end to end through the public host API: H2D + D2H inside the timed region.  `e2e.value` is the contract's
    #      pinned-host number; `e2e.pageable` is the same call on ordinary (pageable) memory, which is what a drop-in
    #      ndarray caller hands over ----
    e2e = None
    if not args.no_e2e:
        eng.set_stream(None)
        ksteps = max(1, min(args.steps, 3))
        fail = C.c_int64(-1)
        # only lower-triangular trapezoids (1024-wide block columns, rows from the diagonal block down) cross PCIe, in arrival
        # waves that overlap the factorisation: csrc/api.cu cholesky_host (option chol_waves)
        tri_bytes = sum(min(n, r0 + 1024) * (min(n, r0 + 1024) - r0) * 8 for r0 in range(0, n, 1024)) if n >= 2048 else n * n * 8

        def e2e_leg(host_src, host_work):
            t_acc = 0.0
            for it in range(ksteps + 1):
                host_work.copy_(host_src)
                barrier()
                t0 = time.perf_counter()
                st = lib.lfb_cholesky_f64(eng.h, C.c_void_p(host_work.data_ptr()), n, n, n, 1, 0, C.byref(fail))
                dt = time.perf_counter() - t0
                if st != 0:
                    raise RuntimeError(f"lfb_cholesky_f64 status {st}")
                if it > 0:
                    t_acc += dt
            return max_over_ranks(t_acc * 1e3)

        host_src = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        host_src.copy_(S)
        host_work = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        torch.cuda.synchronize()
        t_e2e_ms = e2e_leg(host_src, host_work)
        # residual of what came back (first 2048-block of the row-major lower factor)
        Lh = torch.tril(host_work[:2048, :2048]).to(dev)
        e2e_resid = float((Lh @ Lh.t() - S[:2048, :2048]).norm() / S[:2048, :2048].norm())
        e2e = {"value": world * flops * ksteps / (t_e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": tri_bytes, "d2h_bytes_per_step": tri_bytes + 8, "steps": ksteps,
               "ms_per_step": t_e2e_ms / ksteps, "api": "lfb_cholesky_f64 (host view, pinned, in place)", "residual_block": e2e_resid}
        del host_work
        try:
            page_work = torch.empty((n, n), dtype=torch.float64)          # ordinary malloc'ed memory
            t_pg_ms = e2e_leg(host_src, page_work)
            e2e["pageable"] = {"value": world * flops * ksteps / (t_pg_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": t_pg_ms / ksteps,
                               "api": "lfb_cholesky_f64 (host view, pageable, in place)"}
            del page_work
        except Exception as ex:
            e2e["pageable"] = {"error": str(ex)[:120]}
        eng.set_stream(stream.cuda_stream)
        del host_src

    del S, work
    torch.cuda.empty_cache()

    extras = {}
    if not args.no_extras:
        extras = run_extras(args, eng, lib, dev, stream, world, rank, timed, barrier)
        torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle as O
        cn = args.cpu_n
        g = np.random.default_rng(1).uniform(-1, 1, (cn, cn))
        s0 = (g + g.T) / 2 + cn * np.eye(cn)
        t0 = time.perf_counter(); st, _ = O.cholesky(s0); dt = time.perf_counter() - t0
        cpu_baseline = {"value": cn ** 3 / 3.0 / dt / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": "port",
                        "sample": f"oracle restatement of src/cholesky.rs:51-83 on one n={cn} SPD matrix ({dt:.1f} s; the "
                                  f"n={n} workload extrapolates as n^3 to {dt * (n / cn) ** 3 / 60:.1f} min); reference is single-threaded"}

    # The same sharded paths through the single-process multi-device C ABI (csrc/multi.cu): rank 0 drives all `world` devices
    # of the box by itself.  The other ranks are DONE at this point: they leave the process group and exit (a rank parked in
    # an NCCL barrier keeps a spinning kernel on its GPU and a spinning host thread on a core, which cost the single-process
    # leg 4x in the first 8-GPU run, profiles/r2_multi_gpu.md).
    if world > 1:
        barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
        time.sleep(2.0)                      # let the other ranks' processes release their devices
    if rank == 0 and not args.no_extras:
        try:
            extras["cabi_multi"] = run_multi_cabi(world)
        except Exception as ex:
            extras["cabi_multi"] = {"error": str(ex)[:300]}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n),
                       "parallelism": "replicas only (one matrix per GPU; the path does not shard)" if world > 1 else "single GPU",
                       "l2": f"input {n * n * 8 / 2**20:.0f} MiB > 126 MiB L2 (inputs larger than L2; restored by a device copy each step)",
                       "residual_block": resid},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "extras": extras,
        }
        # The N-dependent numbers of the paths that SHARD (batch split / TSQR row split), compact and LAST so that they
        # survive a truncated tail of this line; speed-ups are against the values the N = 1 run of the same box left
        # behind in .bench_n1.json (null if that run did not happen here).
        line["sharded"] = sharded_summary(extras, world)
        print(json.dumps(line), flush=True)


def run_multi_cabi(world):
    """C3 and C4 through `lfb_*_multi_dev_*`: ONE process, `world` devices, NCCL inside the library.  Device-timed
    (start/stop events on every device's stream, max over devices); restore copies are outside the timed region."""
    import numpy as np
    import torch
    from linfa_linalg_b200.dist import MultiEngine, shard_range
    out = {"devices": world}
    m = MultiEngine(n_devices=world)
    out["nccl_ranks"] = m.nccl_ranks
    devs = [torch.device("cuda", i) for i in range(world)]

    def sync_all():
        for d in devs:
            torch.cuda.synchronize(d)

    def timed_multi(restore, call, steps, warmup):
        tot = 0.0
        for it in range(warmup + steps):
            restore()
            sync_all()
            m.time_begin()
            call()
            ms = m.time_end()
            if it >= warmup:
                tot += ms
        return tot / steps
    # ---- C4: 4,194,304 x 256 f64, rows sharded ----
    rows_total, cols = 4194304, 256
    spans = [shard_range(rows_total, world, i) for i in range(world)]
    T0, Tw, Rs, Ds = [], [], [], []
    for i, d in enumerate(devs):
        g = torch.Generator(device=d).manual_seed(0x1F2E3D4C + 4 + i)
        T0.append(torch.rand((cols, spans[i][1] - spans[i][0]), dtype=torch.float64, device=d, generator=g).mul_(2).sub_(1))
        Tw.append(torch.empty_like(T0[-1]))
        Rs.append(torch.zeros((cols, cols), dtype=torch.float64, device=d))
        Ds.append(torch.zeros(cols, dtype=torch.float64, device=d))

    def restore():
        for a, b in zip(Tw, T0):
            a.copy_(b)
    ms_r = timed_multi(restore, lambda: m.tsqr_r_dev(Tw, Rs), 2, 1)
    gram = sum((t @ t.t()).to(devs[0]) for t in T0)
    Rm = Rs[0].t()
    out["tsqr_r_ms"] = ms_r
    out["tsqr_r_check"] = float((Rm.t() @ Rm - gram).norm() / gram.norm())
    ms_q = timed_multi(restore, lambda: m.qr_tsqr_dev(Tw, Ds, Rs), 2, 1)
    Rm = Rs[0].t()
    vn = 0.0
    for i, t in enumerate(Tw):
        low = t.clone()
        if i == 0:
            low[:, :cols] = torch.triu(low[:, :cols])
        ss = (low * low).sum(dim=1).to(devs[0])
        vn = ss if i == 0 else vn + ss
    out["tsqr_qr_ms"] = ms_q
    out["tsqr_qr_check"] = {"RtR_vs_AtA": float((Rm.t() @ Rm - gram).norm() / gram.norm()),
                            "reflector_norm_err": float((vn.sqrt() - 1).abs().max()),
                            "diag_vs_r_err": float((Ds[0].abs() - torch.diagonal(Rs[0])).abs().max()),
                            "diag_replicated": all(bool(torch.equal(Ds[0].cpu(), x.cpu())) for x in Ds)}
    del T0, Tw, gram, low
    for d in devs:
        with torch.cuda.device(d):
            torch.cuda.empty_cache()
    # ---- C3: 262144 x (32 x 32) f32, batch sharded ----
    B = 262144
    bs = [shard_range(B, world, i) for i in range(world)]
    M0, M, Dg = [], [], []
    for i, d in enumerate(devs):
        g = torch.Generator(device=d).manual_seed(0x1F2E3D4C + 3 + i)
        M0.append(torch.rand((bs[i][1] - bs[i][0], 32, 32), dtype=torch.float32, device=d, generator=g).mul_(2).sub_(1))
        M.append(torch.empty_like(M0[-1]))
        Dg.append(torch.zeros((bs[i][1] - bs[i][0], 32), dtype=torch.float32, device=d))

    def restore_b():
        for a, b in zip(M, M0):
            a.copy_(b)
    ms_b = timed_multi(restore_b, lambda: m.qr_batched_dev_f32(M, Dg), 10, 3)
    import oracle as O
    ref = M0[-1][:1024].cpu().numpy().copy()
    dref = O.qr_batched(ref)
    out["batched_ms"] = ms_b
    out["batched_mps"] = B / (ms_b * 1e-3)
    out["batched_check"] = float(max(np.max(np.abs(M[-1][:1024].cpu().numpy() - ref)), np.max(np.abs(Dg[-1][:1024].cpu().numpy() - dref))))
    out["check_ok"] = (out["tsqr_r_check"] <= 1e-12 and out["tsqr_qr_check"]["RtR_vs_AtA"] <= 1e-12 and out["tsqr_qr_check"]["reflector_norm_err"] <= 1e-12
                       and out["tsqr_qr_check"]["diag_replicated"] and out["batched_check"] <= 16 * 32 * 1.2e-7 * 32 ** 0.5)
    out["launches"] = m.launch_count
    m.close()
    return out


N1_FILE = os.path.join(ROOT, ".bench_n1.json")


def sharded_summary(extras, world):
    def g(key, field):
        v = extras.get(key, {})
        return v.get(field) if isinstance(v, dict) else None
    cur = {"batched_mps": g("batched_qr_f32", "matrices_per_s"), "batched_chol_mps": g("batched_chol_f32", "matrices_per_s"),
           "tsqr_r_ms": g("tsqr_f64", "ms_per_step"), "tsqr_qr_ms": g("tsqr_qr_f64", "ms_per_step"),
           # the single-process C-ABI route (lfb_*_multi_dev_*), same workloads
           "cabi_batched_mps": g("cabi_multi", "batched_mps"), "cabi_tsqr_r_ms": g("cabi_multi", "tsqr_r_ms"),
           "cabi_tsqr_qr_ms": g("cabi_multi", "tsqr_qr_ms")}
    out = {"n": world}
    out.update({k: (round(v, 4) if isinstance(v, float) and v < 1e4 else (float(f"{v:.4g}") if v is not None else None)) for k, v in cur.items()})
    out["checks_ok"] = all(bool(g(k, "check_ok")) for k in ("batched_qr_f32", "batched_chol_f32", "tsqr_f64", "tsqr_qr_f64", "cabi_multi") if k in extras)
    out["cabi_nccl_ranks"] = g("cabi_multi", "nccl_ranks")
    if world == 1:
        try:
            json.dump(cur, open(N1_FILE, "w"))
        except OSError:
            pass
        out["speedup_vs_1gpu"] = None
        return out
    sp = None
    try:
        one = json.load(open(N1_FILE))
        sp = {}
        for k, v in cur.items():
            if v and one.get(k):
                sp[k.replace("_ms", "").replace("_mps", "")] = round((v / one[k]) if k.endswith("mps") else (one[k] / v), 3)
    except (OSError, ValueError):
        sp = None
    out["speedup_vs_1gpu"] = sp
    return out


def run_extras(args, eng, lib, dev, stream, world, rank, timed, barrier):
    """C2' (QR 16384^2 f64, replicas), C3 (batched QR, batch-sharded), C4 (TSQR, row-sharded), C5 (eigh / SVD phase 1,
    eigh end to end)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    out = {}
    # ---- C1 (BASELINE configs[0]): QR of a 512 x 512 random f64 matrix through the QR-trait mirror, GPU beside the CPU port ----
    if rank == 0:
        try:
            import oracle as O
            import linfa_linalg_b200 as L
            n1 = 512
            a0 = np.random.default_rng(0x1F2E3D4C).uniform(-1, 1, (n1, n1))
            ref = a0.copy()
            t0 = time.perf_counter(); dref = O.qr(ref); t_cpu = time.perf_counter() - t0     # qr.rs:32-44 on one core
            eng.set_stream(None)
            t_gpu = []
            for _ in range(6):
                a = a0.copy()
                torch.cuda.synchronize()
                t0 = time.perf_counter(); dec = L.qr_into(a, eng=eng); t_gpu.append(time.perf_counter() - t0)
            eng.set_stream(stream.cuda_stream)
            c1_err = float(max(np.max(np.abs(a - ref)), np.max(np.abs(dec.diag - dref))))
            d_a = torch.from_numpy(np.ascontiguousarray(a0.T)).to(dev)
            d_w = torch.empty_like(d_a); d_d = torch.empty(n1, dtype=torch.float64, device=dev)

            def c1_step():
                d_w.copy_(d_a)
                lib.lfb_qr_dev_f64(eng.h, C.c_void_p(d_w.data_ptr()), n1, n1, n1, C.c_void_p(d_d.data_ptr()))
            ms_dev = timed(c1_step, 20, 3) / 20 if world == 1 else None
            fl1 = 4.0 / 3.0 * n1 ** 3
            t_best = min(t_gpu[1:])
            out["c1_qr512_f64"] = {"workload": "QR 512x512 f64 via qr_into (C1, qr.rs:29-63)", "gpu_e2e_ms": t_best * 1e3,
                                   "gpu_e2e_gflops": fl1 / t_best / 1e9, "gpu_device_ms": ms_dev,
                                   "cpu_port_ms": t_cpu * 1e3, "cpu_port_gflops": fl1 / t_cpu / 1e9, "cpu_cores": 1,
                                   "e2e_speedup_vs_cpu_port": t_cpu / t_best,
                                   "check": {"max_abs_err_vs_oracle": c1_err}, "check_ok": c1_err <= 16 * n1 * 2.3e-16 * float(np.linalg.norm(a0))}
        except Exception as ex:
            out["c1_qr512_f64"] = {"error": str(ex)[:200]}
    # ---- C2': blocked Householder QR, n = 16384 f64 ----
    try:
        n = args.n
        gen = torch.Generator(device=dev).manual_seed(0x1F2E3D4C + 2 + rank)
        A0 = torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen).mul_(2).sub_(1)
        A = torch.empty_like(A0)
        diag = torch.empty(n, dtype=torch.float64, device=dev)

        def qr_step():
            A.copy_(A0)
            st = lib.lfb_qr_dev_f64(eng.h, C.c_void_p(A.data_ptr()), n, n, n, C.c_void_p(diag.data_ptr()))
            if st != 0:
                raise RuntimeError(f"lfb_qr_dev_f64 status {st}")
        ms = timed(qr_step, 2, 1)
        fl = 4.0 / 3.0 * n ** 3
        # correctness of the timed result (size-independent): the leading k x k block of R^T R equals that of A^T A
        # (R upper triangular), diag(R) = |diag| and every reflector has unit norm (householder.rs:23)
        k = min(2048, n)
        Rk = torch.triu(A[:k, :k].t(), 1) + torch.diag(diag[:k].abs())
        AtA = A0[:k, :] @ A0[:k, :].t()
        rtr_err = float((Rk.t() @ Rk - AtA).norm() / AtA.norm())
        vnorm = torch.triu(A[:k, :]).pow(2).sum(dim=1).sqrt()          # tensor row c = column c of the factor; triu keeps rows >= c
        vn_err = float((vnorm - 1).abs().max())
        ok = rtr_err <= 8 * n * 2.3e-16 and vn_err <= 1e-13 and bool(torch.isfinite(diag).all())
        out["qr_f64"] = {"workload": f"QR {n}x{n} f64 (C2')", "gflops": world * fl * 2 / (ms * 1e-3) / 1e9, "ms_per_step": ms / 2,
                         "frac_of_nominal_fp64": fl * 2 / (ms * 1e-3) / 1e12 / FP64_NOMINAL_TFLOPS,
                         "check": {"RtR_vs_AtA_block2048": rtr_err, "reflector_norm_err": vn_err}, "check_ok": ok}
        # e2e through the host API (lfb_qr_f64, row-major pinned view, in place): upload, transpose, factor, download
        if not args.no_e2e and world == 1:
            hq = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
            hd = np.zeros(n)
            hq.copy_(A0)
            eng.set_stream(None)
            t_q = []
            for it in range(2):
                hq.copy_(A0)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                st = lib.lfb_qr_f64(eng.h, C.c_void_p(hq.data_ptr()), n, n, n, 1, C.c_void_p(hd.ctypes.data))
                t_q.append(time.perf_counter() - t0)
                if st != 0:
                    raise RuntimeError(f"lfb_qr_f64 status {st}")
            eng.set_stream(stream.cuda_stream)
            out["qr_f64"]["e2e"] = {"gflops": fl / t_q[-1] / 1e9, "ms": t_q[-1] * 1e3, "h2d_bytes": n * n * 8, "d2h_bytes": n * n * 8 + n * 8,
                                    "api": "lfb_qr_f64 (host view, pinned, in place)"}
            del hq
        del A0, A
        torch.cuda.empty_cache()
    except Exception as ex:  # keep the headline line alive
        out["qr_f64"] = {"error": str(ex)[:200]}
    # ---- f32 engine: Cholesky / QR 16384^2 f32, trailing updates on tcgen05 3xTF32 (csrc/gemm_tf32.cu); replicas ----
    try:
        n = args.n
        pk = C.c_double(0)
        lib.lfb_microbench_fp64(eng.h, 2, C.byref(pk))
        ffma_peak = pk.value / 1e3
        gen = torch.Generator(device=dev).manual_seed(0x1F2E3D4C + 6 + rank)
        F0 = torch.rand((n, n), dtype=torch.float32, device=dev, generator=gen).mul_(2).sub_(1)
        Fw = torch.empty_like(F0)
        fd = torch.empty(n, dtype=torch.float32, device=dev)
        finfo = torch.zeros(1, dtype=torch.int64, device=dev)

        def qrf_step():
            Fw.copy_(F0)
            st = lib.lfb_qr_dev_f32(eng.h, C.c_void_p(Fw.data_ptr()), n, n, n, C.c_void_p(fd.data_ptr()))
            if st != 0:
                raise RuntimeError(f"lfb_qr_dev_f32 status {st}")
        ms = timed(qrf_step, 2, 1) / 2
        k = min(2048, n)
        Rk = (torch.triu(Fw[:k, :k].t(), 1) + torch.diag(fd[:k].abs())).double()
        AtA = F0[:k, :].double() @ F0[:k, :].double().t()
        rtr = float((Rk.t() @ Rk - AtA).norm() / AtA.norm())
        fl = 4.0 / 3.0 * n ** 3
        out["qr_f32"] = {"workload": f"QR {n}x{n} f32", "ms_per_step": ms, "gflops": world * fl / (ms * 1e-3) / 1e9,
                         "ffma_peak_tflops": ffma_peak, "frac_of_ffma_peak": fl / (ms * 1e-3) / 1e12 / ffma_peak,
                         "trailing_update": "tcgen05.mma.kind::tf32, 3xTF32 split, TMEM accumulator promoted every 64 k (csrc/gemm_tf32.cu)",
                         "check": {"RtR_vs_AtA_block2048": rtr}, "check_ok": rtr <= 8 * n * 1.2e-7}
        F0.copy_((F0 + F0.t()) * 0.5)
        F0.diagonal().add_(float(n))

        def cholf_step():
            Fw.copy_(F0)
            st = lib.lfb_cholesky_dev_f32(eng.h, C.c_void_p(Fw.data_ptr()), n, n, 0, C.c_void_p(finfo.data_ptr()))
            if st != 0:
                raise RuntimeError(f"lfb_cholesky_dev_f32 status {st}")
        ms = timed(cholf_step, 3, 1) / 3
        Lf = torch.triu(Fw[:k, :k]).t().double()
        Sd = F0[:k, :k].double()
        cres = float((Lf @ Lf.t() - Sd).norm() / Sd.norm())
        fl = n ** 3 / 3.0
        out["chol_f32"] = {"workload": f"Cholesky {n}x{n} f32", "ms_per_step": ms, "gflops": world * fl / (ms * 1e-3) / 1e9,
                           "ffma_peak_tflops": ffma_peak, "frac_of_ffma_peak": fl / (ms * 1e-3) / 1e12 / ffma_peak,
                           "check": {"residual_block2048": cres}, "check_ok": cres <= 8 * n * 1.2e-7 and int(finfo.item()) == 0}
        # the trailing-update kernel alone on the SYRK shape of the factorisation (TN, K = 512)
        kk = 512
        Pa = torch.rand((n, kk), dtype=torch.float32, device=dev, generator=gen).sub_(0.5)
        Cc = torch.zeros((n, n), dtype=torch.float32, device=dev)

        def gemm_step():
            lib.lfb_gemm_dev_f32(eng.h, 1, 0, n, n, kk, 1.0, C.c_void_p(Pa.data_ptr()), kk, C.c_void_p(Pa.data_ptr()), kk, 0.0,
                                 C.c_void_p(Cc.data_ptr()), n)
        ms_g = timed(gemm_step, 5, 2) / 5
        tf = 2.0 * n * n * kk / (ms_g * 1e-3) / 1e12
        ref = Pa[:256].double() @ Pa[:256].double().t()
        gerr = float((Cc[:256, :256].double() - ref).abs().max())
        out["sgemm_tc"] = {"workload": f"f32 TN GEMM {n}x{n}x{kk} (3xTF32 on tcgen05)", "tflops": tf, "frac_of_ffma_peak": tf / ffma_peak,
                           "check": {"max_abs_err_vs_f64": gerr}, "check_ok": gerr <= 2e-6 * kk}
        del F0, Fw, Pa, Cc
        torch.cuda.empty_cache()
    except Exception as ex:
        out["f32_engine"] = {"error": str(ex)[:200]}
    # ---- C3: batched 32x32 f32 QR, batch-sharded (strong scaling) ----
    try:
        B = 262144
        per = B // world
        gen = torch.Generator(device=dev).manual_seed(0x1F2E3D4C + 3 + rank)
        M0 = torch.rand((per, 32, 32), dtype=torch.float32, device=dev, generator=gen).mul_(2).sub_(1)
        M = torch.empty_like(M0)
        d = torch.empty((per, 32), dtype=torch.float32, device=dev)

        def bq_step():
            M.copy_(M0)
            st = lib.lfb_qr_batched_dev_f32(eng.h, C.c_void_p(M.data_ptr()), per, 32, 32, C.c_void_p(d.data_ptr()))
            if st != 0:
                raise RuntimeError(f"lfb_qr_batched_dev_f32 status {st}")

        def copy_only():
            M.copy_(M0)
        ms_all = timed(bq_step, 20, 5)
        ms_copy = timed(copy_only, 20, 5)
        ms_k = max(ms_all - ms_copy, 1e-6)
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        gbs = per * 8320 * 20 / (ms_k * 1e-3) / 1e9
        # correctness of the timed result: 4096 sampled matrices of this rank's shard against the CPU oracle (qr.rs:32-44
        # per matrix), elementwise + the sign bits of diag (householder.rs:50)
        import oracle as O
        bq_step()                                   # the last timed call was the restore copy: factor once more for the check
        idx = torch.randperm(per, device=dev, generator=gen)[:4096].sort().values
        ref = M0[idx].cpu().numpy().copy()
        dref = O.qr_batched(ref)
        got, dgot = M[idx].cpu().numpy(), d[idx].cpu().numpy()
        # A pivot's sign is the sign of the column head at that step (householder.rs:16); a head within rounding of zero may
        # come out with the other sign under a different summation order -- about one pivot in a million on this data --
        # and flips its row of R and its reflector.  Matrices whose sign bits agree are compared elementwise; the rare others
        # must still agree in |diag| (the observable R diagonal, qr.rs:96).
        same = np.all(np.signbit(dgot) == np.signbit(dref), axis=1)
        berr = float(max(np.max(np.abs(got[same] - ref[same])), np.max(np.abs(dgot[same] - dref[same]))))
        flipped = int((~same).sum())
        ferr = float(np.max(np.abs(np.abs(dgot[~same]) - np.abs(dref[~same])))) if flipped else 0.0
        sign_ok = flipped <= 4 and ferr <= 1e-3
        # elementwise tolerance per matrix: 16 n eps sqrt(n) for a well-conditioned sample, widened in proportion to the spread of
        # |diag(R)| (a lower bound of cond_2): among the 32768 matrices sampled over 8 ranks a few have a pivot ~1e-3 of the
        # largest, and the forward error of ANY backward-stable QR -- the oracle's too -- grows with that ratio
        base_tol = 16 * 32 * 1.2e-7 * 32 ** 0.5
        kap = np.abs(dref).max(axis=1) / np.maximum(np.abs(dref).min(axis=1), 1e-30)
        e_b = np.maximum(np.abs(got - ref).reshape(len(ref), -1).max(axis=1), np.abs(dgot - dref).max(axis=1))
        tol_b = base_tol * np.maximum(1.0, kap / 64.0)
        worst = float(np.max(e_b[same] / tol_b[same])) if same.any() else 0.0
        ok_t = torch.tensor([1 if (worst <= 1.0 and sign_ok) else 0], device=dev)
        if world > 1:
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        out["batched_qr_f32"] = {"workload": f"{B} x (32x32) f32, {per} per GPU (C3)", "matrices_per_s": B * 20 / (ms_k * 1e-3),
                                 "ms_per_step": ms_k / 20, "kernel_GBps_per_gpu": gbs, "hbm_peak_GBps": hbm,
                                 "frac_of_hbm": gbs / hbm, "scaling": "strong", "note": "restore copy timed separately and subtracted",
                                 "check": {"sampled": 4096, "max_abs_err_vs_oracle": berr, "worst_err_over_conditioned_tol": worst,
                                           "matrices_with_a_flipped_near_zero_pivot": flipped},
                                 "check_ok": bool(ok_t.item())}
        # batched Cholesky on the same shard (north star: "batched small-matrix QR/Cholesky is split by batch")
        M0.copy_(torch.bmm(M0, M0.transpose(1, 2)))
        M0.diagonal(dim1=1, dim2=2).add_(32.0)
        fl_t = torch.empty(per, dtype=torch.int32, device=dev)

        def bc_step():
            M.copy_(M0)
            st = lib.lfb_cholesky_batched_dev_f32(eng.h, C.c_void_p(M.data_ptr()), per, 32, 0, C.c_void_p(fl_t.data_ptr()))
            if st != 0:
                raise RuntimeError(f"lfb_cholesky_batched_dev_f32 status {st}")
        ms_c = max(timed(bc_step, 20, 5) - ms_copy, 1e-6)
        refc = M0[idx].cpu().numpy().copy()
        fm, _ = O.cholesky_batched(refc, False)
        gotc = M[idx].cpu().numpy()
        cerr = float(np.max(np.abs(np.tril(gotc) - np.tril(refc))))
        okc = fm == -1 and int(fl_t.max().item()) < 0 and cerr <= 16 * 32 * 1.2e-7 * 32 ** 0.5 * float(M0.abs().max())
        gbc = per * (4096 + 4096) * 20 / (ms_c * 1e-3) / 1e9
        out["batched_chol_f32"] = {"workload": f"{B} x (32x32) f32 SPD, {per} per GPU", "matrices_per_s": B * 20 / (ms_c * 1e-3),
                                   "ms_per_step": ms_c / 20, "kernel_GBps_per_gpu": gbc, "frac_of_hbm": gbc / hbm, "scaling": "strong",
                                   "check": {"sampled": 4096, "max_abs_err_vs_oracle": cerr}, "check_ok": bool(okc)}
        del M0, M
        torch.cuda.empty_cache()
    except Exception as ex:
        out["batched_qr_f32"] = {"error": str(ex)[:200]}
    # ---- C4: TSQR, rows sharded across ranks, R factors all-gathered over NCCL ----
    try:
        rows_total, cols = 4194304, 256
        rows = rows_total // world
        gen = torch.Generator(device=dev).manual_seed(0x1F2E3D4C + 4 + rank)
        # column-major rows x cols block == row-major (cols, rows) tensor
        T0 = torch.rand((cols, rows), dtype=torch.float64, device=dev, generator=gen).mul_(2).sub_(1)
        Tw = torch.empty_like(T0)
        from linfa_linalg_b200 import dist as D
        local_r = D.gpu_local_r(eng, Tw, rows, cols)
        final_r = D.gpu_final_r(eng, cols)

        rres = {}

        def tsqr_step():
            Tw.copy_(T0)
            # local R per rank -> one NCCL all_gather of the 256x256 factors -> R of the stack (replicated)
            rres["r"] = D.tsqr_r(local_r, final_r, cols)
        ms = timed(tsqr_step, 2, 2)
        fl = 2.0 * rows_total * cols * cols - 2.0 / 3.0 * cols ** 3
        # correctness of the timed result: R^T R = A^T A (all-reduced Gram matrix of the row shards), ||R||_F = ||A||_F,
        # diag(R) >= 0 and strict lower triangle exactly zero (qr.rs:93-96)
        Rm = rres["r"].t().clone()                            # math R (the tensor is its column-major storage; the callable reuses it)
        gram = T0 @ T0.t()
        if world > 1:
            dist.all_reduce(gram)
        g_err = float((Rm.t() @ Rm - gram).norm() / gram.norm())
        nrm_err = float(abs(Rm.norm() / gram.diagonal().sum().sqrt() - 1))
        struct_ok = bool((Rm.diagonal() >= 0).all()) and bool((torch.tril(Rm, -1) == 0).all())
        out["tsqr_f64"] = {"workload": f"TSQR {rows_total}x{cols} f64, {rows} rows per GPU (C4)", "gflops": fl * 2 / (ms * 1e-3) / 1e9,
                           "ms_per_step": ms / 2, "scaling": "strong", "exchange": "NCCL all_gather of 256x256 R per rank" if world > 1 else "none",
                           "check": {"RtR_vs_AtA": g_err, "normR_over_normA_minus_1": nrm_err, "diag_nonneg_lower_zero": struct_ok},
                           "check_ok": g_err <= 1e-12 and nrm_err <= 1e-13 and struct_ok}
        out["tsqr_f64"]["leaf"] = "Cholesky-QR leaf (Gram GEMM + 256x256 Cholesky, cond guard; csrc/cholqr.cu)"
        if world == 1:      # the Householder leaf (round 1's path; still the fallback for ill-conditioned blocks) beside it
            eng.set_option("tsqr_cholqr_cond", 0)
            try:
                ms_h = timed(tsqr_step, 1, 1)
                out["tsqr_f64"]["householder_leaf_ms"] = ms_h
                Rh = rres["r"].t()
                out["tsqr_f64"]["leaf_vs_householder_max_rel_diff"] = float((Rh - Rm).abs().max() / Rm.abs().max())
            finally:
                eng.set_option("tsqr_cholqr_cond", 16)
        del T0, Tw, gram
        torch.cuda.empty_cache()
    except Exception as ex:
        out["tsqr_f64"] = {"error": str(ex)[:200]}
    # ---- C5a / C5b: phase 1 of eigh (tridiagonalisation + Q) and of SVD (bidiagonalisation); replicas ----
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        n = 8192
        gen = torch.Generator(device=dev).manual_seed(0x1F2E3D4C + 5 + rank)
        S0 = torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen).mul_(2).sub_(1)
        S0 = S0.add_(S0.t().clone()).mul_(0.5)
        Sw = torch.empty_like(S0)
        Q = torch.empty_like(S0)
        off = torch.zeros(n, dtype=torch.float64, device=dev)

        def trd_step():
            Sw.copy_(S0)
            st = lib.lfb_sym_tridiagonal_dev_f64(eng.h, C.c_void_p(Sw.data_ptr()), n, n, C.c_void_p(off.data_ptr()))
            if st != 0:
                raise RuntimeError(f"lfb_sym_tridiagonal_dev_f64 status {st}")

        def q_step():
            st = lib.lfb_assemble_q_dev_f64(eng.h, C.c_void_p(Sw.data_ptr()), n, n, n, 1, C.c_void_p(off.data_ptr()), C.c_void_p(Q.data_ptr()), n)
            if st != 0:
                raise RuntimeError(f"lfb_assemble_q_dev_f64 status {st}")
        ms_t = timed(trd_step, 1, 1)
        ms_q = timed(q_step, 1, 1)
        fl = 4.0 / 3.0 * n ** 3
        symv_bytes = 4.0 * n ** 3 / 3.0       # ONE read of the lower triangle of the trailing matrix per column
        us_symv = C.c_double(0)
        lib.lfb_microbench_kernel(eng.h, b"trd_symv", n, 50, C.byref(us_symv))       # the SYMV alone, back to back, at the full size
        # orthogonal-similarity invariants of the timed result: trace(T) = trace(A), ||T||_F = ||A||_F
        td = Sw.diagonal()
        tr_err = float(abs(td.sum() - S0.diagonal().sum()) / S0.norm())
        fro_err = float(abs(torch.sqrt((td * td).sum() + 2 * (off[:n - 1] ** 2).sum()) / S0.norm() - 1))
        out["tridiag_f64"] = {"workload": f"sym_tridiagonal {n}x{n} f64 (C5a phase 1)", "ms": ms_t, "gflops": fl / (ms_t * 1e-3) / 1e9,
                              "check": {"trace_err": tr_err, "fro_norm_err": fro_err}, "check_ok": tr_err <= 1e-12 and fro_err <= 1e-12,
                              "symv_GBps_lower_bound": symv_bytes / (ms_t * 1e-3) / 1e9, "hbm_peak_GBps": hbm,
                              "frac_of_hbm_lower_bound": symv_bytes / (ms_t * 1e-3) / 1e9 / hbm,
                              "symv_kernel_at_n": {"us": us_symv.value, "GBps": 4.0 * n * n / max(us_symv.value, 1e-9) / 1e3,
                                                   "frac_of_hbm": 4.0 * n * n / max(us_symv.value, 1e-9) / 1e3 / hbm},
                              "generate_q_ms": ms_q, "generate_q_gflops": fl / (ms_q * 1e-3) / 1e9}
        # eigh end to end (eigh.rs:10-129): scale, tridiagonalise, Q, implicit-QR Givens phase (host recurrence,
        # rotations applied on the device 16 sweeps per pass); eigenvalues land in host memory
        vals = np.zeros(n)

        def eigh_step():
            Sw.copy_(S0)
            st = lib.lfb_eigh_dev_f64(eng.h, C.c_void_p(Sw.data_ptr()), n, n, C.c_void_p(vals.ctypes.data), C.c_void_p(Q.data_ptr()), n)
            if st != 0:
                raise RuntimeError(f"lfb_eigh_dev_f64 status {st}")
        barrier()
        t0 = time.perf_counter(); eigh_step(); torch.cuda.synchronize(); t_eigh = time.perf_counter() - t0
        Qt = Q.t()                                  # torch sees the column-major Q transposed
        resid = float((S0 @ Qt - Qt * torch.from_numpy(vals).to(dev)[None, :]).norm() / S0.norm())
        out["eigh_f64"] = {"workload": f"eigh {n}x{n} f64, eigenvalues + eigenvectors (C5a end to end)", "ms": t_eigh * 1e3,
                           "relative_residual": resid, "check_ok": resid <= 64 * n * 2.3e-16, "timing": "wall clock around one call (host recurrence inside)"}
        del S0, Sw, Q, Qt
        torch.cuda.empty_cache()
        m2, n2 = 16384, 4096
        B0 = torch.rand((n2, m2), dtype=torch.float64, device=dev, generator=gen).mul_(2).sub_(1)   # column-major m2 x n2
        Bw = torch.empty_like(B0)
        dd = torch.zeros(n2, dtype=torch.float64, device=dev)
        ee = torch.zeros(n2, dtype=torch.float64, device=dev)

        def bd_step():
            Bw.copy_(B0)
            st = lib.lfb_bidiagonal_dev_f64(eng.h, C.c_void_p(Bw.data_ptr()), m2, n2, m2, C.c_void_p(dd.data_ptr()), C.c_void_p(ee.data_ptr()))
            if st != 0:
                raise RuntimeError(f"lfb_bidiagonal_dev_f64 status {st}")
        ms_b = timed(bd_step, 1, 0)
        flb = 4.0 * m2 * n2 * n2 - 4.0 / 3.0 * n2 ** 3
        gemv_bytes = 16.0 * (m2 * n2 * n2 / 2.0 - n2 ** 3 / 6.0)   # two streaming passes per column/row pair
        us_n, us_t = C.c_double(0), C.c_double(0)
        lib.lfb_microbench_kernel(eng.h, b"bd_gemv_n", n2, 50, C.byref(us_n))     # the GEMVs alone at the full 16384 x 4096 size
        lib.lfb_microbench_kernel(eng.h, b"bd_gemv_t", n2, 50, C.byref(us_t))
        full = 8.0 * m2 * n2
        bfro = float(abs(torch.sqrt((dd * dd).sum() + (ee[:n2 - 1] ** 2).sum()) / B0.norm() - 1))      # ||B||_F = ||A||_F
        out["bidiag_f64"] = {"workload": f"bidiagonal {m2}x{n2} f64 (C5b phase 1, blocked: deferred rank-1 updates)", "ms": ms_b,
                             "check": {"fro_norm_err": bfro}, "check_ok": bfro <= 1e-12,
                             "gflops": flb / (ms_b * 1e-3) / 1e9, "gemv_GBps_lower_bound": gemv_bytes / (ms_b * 1e-3) / 1e9,
                             "hbm_peak_GBps": hbm, "frac_of_hbm_lower_bound": gemv_bytes / (ms_b * 1e-3) / 1e9 / hbm,
                             "gemv_kernels_at_full_size": {"n_us": us_n.value, "n_GBps": full / max(us_n.value, 1e-9) / 1e3,
                                                           "t_us": us_t.value, "t_GBps": full / max(us_t.value, 1e-9) / 1e3}}
        # SVD end to end (svd.rs:17-221): scale, bidiagonalise, U and V, Golub-Kahan Givens phase
        Ud = torch.empty((n2, m2), dtype=torch.float64, device=dev)       # column-major m2 x n2
        Vd = torch.empty((n2, n2), dtype=torch.float64, device=dev)       # column-major n2 x n2, V = Vt^T
        sv = np.zeros(n2)

        def svd_step():
            Bw.copy_(B0)
            st = lib.lfb_svd_dev_f64(eng.h, C.c_void_p(Bw.data_ptr()), m2, n2, m2, C.c_void_p(sv.ctypes.data),
                                     C.c_void_p(Ud.data_ptr()), m2, C.c_void_p(Vd.data_ptr()), n2)
            if st != 0:
                raise RuntimeError(f"lfb_svd_dev_f64 status {st}")
        barrier()
        t0 = time.perf_counter(); svd_step(); torch.cuda.synchronize(); t_svd = time.perf_counter() - t0
        # torch sees the column-major buffers transposed: Ud = U^T (n2 x m2), Vd = V^T = Vt (n2 x n2); B0 = A^T
        rec = (Vd.t() * torch.from_numpy(sv).to(dev)[None, :]) @ Ud                  # (U S Vt)^T = V S U^T
        resid = float((rec - B0).norm() / B0.norm())
        out["svd_f64"] = {"workload": f"svd {m2}x{n2} f64 with U and Vt (C5b end to end)", "ms": t_svd * 1e3,
                          "relative_residual": resid, "check_ok": resid <= 64 * m2 * 2.3e-16, "timing": "wall clock around one call (host recurrence inside)"}
        del B0, Bw, Ud, Vd, rec
        torch.cuda.empty_cache()
    except Exception as ex:
        out["tridiag_bidiag"] = {"error": str(ex)[:200]}
    # ---- C4 as a drop-in for qr_into: TSQR + Householder reconstruction -> the reference's compact
    #      factor, rows sharded across ranks; all_gather of R, broadcast of U' and diag.  Last: it is the newest path. ----
    try:
        rows_total, cols = 4194304, 256
        rows = rows_total // world
        from linfa_linalg_b200 import dist as D
        gen = torch.Generator(device=dev).manual_seed(0x1F2E3D4C + 4 + rank)
        T0 = torch.rand((cols, rows), dtype=torch.float64, device=dev, generator=gen).mul_(2).sub_(1)
        Tw = torch.empty_like(T0)
        ops = D.GpuTsqrOps(eng)
        res = {}

        def tsqr_qr_step():
            Tw.copy_(T0)
            res["diag"], res["r"] = D.tsqr_qr(Tw, ops, cols)
        ms = timed(tsqr_qr_step, 4, 2)       # (the route has host synchronisations on every rank: two steps were at the mercy of one hiccup)
        fl = 2.0 * rows_total * cols * cols - 2.0 / 3.0 * cols ** 3
        # self-check on the device: every reflector has unit norm (householder.rs:23), |diag| = diag(R), ||R||_F = ||A||_F
        low = Tw.clone()
        if rank == 0:
            low[:, :cols] = torch.triu(low[:, :cols])          # torch sees the block transposed: keep v (on and below the diagonal)
        ss = (low * low).sum(dim=1)
        a2 = (T0 * T0).sum().reshape(1)
        if world > 1:
            dist.all_reduce(ss)
            dist.all_reduce(a2)
        r = res["r"]
        out["tsqr_qr_f64"] = {"workload": f"qr_into via TSQR + Householder reconstruction {rows_total}x{cols} f64, {rows} rows per GPU",
                              "gflops_qr_equiv": fl * 4 / (ms * 1e-3) / 1e9, "ms_per_step": ms / 4, "scaling": "strong",
                              "reflector_norm_err": float((ss.sqrt() - 1).abs().max()),
                              "diag_vs_r_err": float((res["diag"].abs() - torch.diagonal(r)).abs().max()),
                              "normR_over_normA_minus_1": float(r.norm() / a2.sqrt()[0] - 1),
                              "check_ok": float((ss.sqrt() - 1).abs().max()) <= 1e-12 and float((res["diag"].abs() - torch.diagonal(r)).abs().max()) <= 1e-9
                              and abs(float(r.norm() / a2.sqrt()[0] - 1)) <= 1e-12,
                              "exchange": "all_gather of R + broadcast of U' and diag" if world > 1 else "none"}
        del T0, Tw, low
        torch.cuda.empty_cache()
    except Exception as ex:
        out["tsqr_qr_f64"] = {"error": str(ex)[:200]}
    return out


if __name__ == "__main__":
    main()
