#!/usr/bin/env python
"""bench.py -- MODWT forward + inverse throughput (GSamples/s) on B200, per BASELINE.json.

A "step" is one pass of the hot path over one batch of synthetic input: multi-level MODWT decompose
(forward) followed by reconstruct (inverse) of every signal of the workload.  Default workload (N=1):
BASELINE.json configs[1], the extensions batch-facade shape 4096 x 4096 fp64 signals, db4, J=4, PERIODIC.

  value     whole-job GSamples/s (B*N input samples per fwd+inv step), inputs resident in HBM
  e2e       same metric through the public host-buffer API (BatchMODWT.multiLevelAoS + inverseMultiLevelAoS
            -> C ABI with HOST pointers): H2D of the signals, D2H of all coefficients, H2D of the coefficients
            and D2H of the reconstruction are all inside the timed region
  roofline  dominant kernel (fused analysis) against the measured HBM copy peak: algorithmic bytes
            (24 B/sample/level) per launch / its CUDA-event duration
  cpu_baseline  the oracle's C restatement of the reference's dense loops (kind "port"; the reference is Java and
            no JVM exists on the box) on all host cores, same workload (bounded sample) = the reference's
            structured-concurrency baseline; `variants` adds SURVEY 8(d)'s one-thread core scalar and SoA batch lanes

Multi-GPU (torchrun, one rank per GPU): the batch shards by signal with no communication (weak scaling: every
rank runs the full per-GPU workload); `--workload span` runs one long coif5 signal span-sharded with NCCL halo
exchange.  `--impl reference` times the CPU restatement only (rank 0; other ranks exit).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GSamples/s MODWT fwd+inv fp64 (J levels)"
UNIT = "GSamples/s"
S = 1.0 / math.sqrt(2.0)

WORKLOADS = {
    # name: (wavelet, batch, n, levels, boundary mode)
    "batch4096x4096_db4_J4": ("db4", 4096, 4096, 4, 0),
    "batch4096x4096_haar_J4": ("haar", 4096, 4096, 4, 0),
    "batch16x4096_db4_J4": ("db4", 16, 4096, 4, 0),
    "batch1024x65536_sym8_J8": ("sym8", 1024, 65536, 8, 0),
    "single2p28_coif5_J10": ("coif5", 1, 1 << 28, 10, 0),
    "batch256x1M_db8_J6": ("db8", 256, 1 << 20, 6, 0),
}
DEFAULT = "batch4096x4096_db4_J4"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_near_gpu(props):
    """Multi-rank runs: keep this process (and the pinned host buffers it first-touches) on the NUMA node the GPU hangs
    off, so eight ranks do not push their PCIe traffic across the socket link.  Best effort: silently does nothing when
    sysfs does not say (containers, single-node hosts)."""
    try:
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_port(workload, threads, budget_rows=None, inverse=True):
    """Times the CPU restatement of the reference's dense loops on a bounded sample of the workload."""
    import numpy as np
    from oracle import cref
    from oracle.wavelets import filters
    wname, b, n, levels, mode = WORKLOADS[workload]
    h, g, wid = filters(wname)
    macs_per_sample = 2 * sum((len(h) - 1) * (1 << j) + 1 for j in range(levels)) * (2 if inverse else 1)
    # ~1.5e9 dense MAC/s/core: size the sample for roughly 10 s of wall time
    target = 10.0 * 1.2e9 * max(threads, 1)
    n_s = min(n, 1 << 20)
    rows = max(1, min(b, int(target / (macs_per_sample * n_s)))) if budget_rows is None else budget_rows
    rows = max(rows, min(b, threads))
    x = np.random.default_rng(42).standard_normal((rows, n_s))
    cref.lib()
    t0 = time.perf_counter()
    cref.batch_fwd_inv(x, h, g, levels, mode, wid, dense=True, threads=threads, inverse=inverse)
    dt = time.perf_counter() - t0
    return {"value": rows * n_s / dt * 1e-9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{rows} x {n_s} of the {b} x {n} workload, {wname} J={levels}, dense upsampled taps as the "
                      f"reference (Sum_j 2*L_j MACs/sample/direction), {dt:.2f} s, pthread work queue over signals",
            "seconds": dt}


def cpu_variants(workload):
    """SURVEY 8(d)'s other two CPU baselines on one host thread, ~2 s each: (1) the core scalar path (one signal after
    the other, dense taps, forward + inverse) and (2) the extensions' SoA batch path (lanes = signals, dense taps,
    forward only as in BatchSIMDMODWT.batchMultiLevelMODWTSoA)."""
    import numpy as np
    from oracle import cref
    from oracle.wavelets import filters
    wname, b, n, levels, mode = WORKLOADS[workload]
    h, g, wid = filters(wname)
    macs_fwd = 2 * sum((len(h) - 1) * (1 << j) + 1 for j in range(levels))
    n_s = min(n, 1 << 16)
    out = []
    rows = max(1, min(b, int(2.0 * 0.4e9 / (2 * macs_fwd * n_s))))
    x = np.random.default_rng(42).standard_normal((rows, n_s))
    t0 = time.perf_counter()
    cref.batch_fwd_inv(x, h, g, levels, mode, wid, dense=True, threads=1, inverse=True)
    dt = time.perf_counter() - t0
    out.append({"name": "core_scalar_1_thread", "value": rows * n_s / dt * 1e-9, "unit": UNIT, "cores": 1,
                "sample": f"{rows} x {n_s}, forward + inverse, {dt:.2f} s"})
    if mode == 0:
        lanes = max(1, min(b, int(2.0 * 0.9e9 / (macs_fwd * n_s))))
        xs = np.ascontiguousarray(np.random.default_rng(43).standard_normal((lanes, n_s)).T).reshape(-1)
        t0 = time.perf_counter()
        cref.batch_soa_decompose(xs, lanes, n_s, h, g, levels)
        dt = time.perf_counter() - t0
        out.append({"name": "soa_batch_lanes_1_thread_forward_only", "value": lanes * n_s / dt * 1e-9,
                    "unit": "GSamples/s (forward only)", "cores": 1,
                    "sample": f"{lanes} lanes x {n_s}, batchMultiLevelMODWTSoA restated (gcc -O2 autovectorised), {dt:.2f} s"})
    return out


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    vals = []
    for _ in range(args.warmup):
        cpu_port(args.workload, threads, budget_rows=max(threads, 8))
    cb = None
    for _ in range(max(1, args.steps)):
        cb = cpu_port(args.workload, threads, budget_rows=args.ref_rows)
        vals.append(cb["value"])
    value = statistics.median(vals)
    wname, b, n, levels, mode = WORKLOADS[args.workload]
    cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "wavelet": wname, "batch": b, "signal_length": n, "levels": levels,
                       "boundary": "PERIODIC"},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement (oracle port) of the reference's scalar loops; the reference is pure Java and "
                    "no JVM exists on the box"}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=sorted(WORKLOADS) + ["span"])
    ap.add_argument("--ref-rows", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "span":
            args.workload = "single2p28_coif5_J10"
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    import vectorwave_b200 as vw

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MODWT engine has no CPU path")
    torch.cuda.set_device(local_rank)
    numa_node = bind_near_gpu(torch.cuda.get_device_properties(local_rank)) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.workload == "span":
        from vectorwave_b200 import sharded
        return sharded.bench_span(args, rank, world, local_rank, METRIC, UNIT, ClockSampler, measured_peaks)

    wname, b, n, levels, mode = WORKLOADS[args.workload]
    eng = vw.Engine.get(local_rank)
    wv = vw.get_wavelet(wname)
    hs, gs = wv.lowPassDecomposition() * S, wv.highPassDecomposition() * S
    dev = torch.device("cuda", local_rank)

    # batch sharding by signal: weak scaling, every rank owns a full per-GPU batch; no data-path collective
    nsets = 3  # rotate input / output sets so no step finds its inputs in L2 (each set is 6x L2 anyway)
    if (levels + 3) * b * n * 8 * nsets > 60e9:
        nsets = 1
    gen = torch.Generator(device=dev)
    gen.manual_seed(42 + rank)
    xs = [torch.randn((b, n), dtype=torch.float64, device=dev, generator=gen) for _ in range(nsets)]
    ws = [torch.empty((levels, b, n), dtype=torch.float64, device=dev) for _ in range(nsets)]
    vs = [torch.empty((b, n), dtype=torch.float64, device=dev) for _ in range(nsets)]
    xr = [torch.empty((b, n), dtype=torch.float64, device=dev) for _ in range(nsets)]

    def step(i):
        k = i % nsets
        eng.forward(xs[k], hs, gs, levels, mode, 0, ws[k], vs[k])
        eng.inverse(ws[k], vs[k], hs, gs, mode, None, 0, out=xr[k])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * b * n / ms_step * 1e-6

    # correctness guard inside the bench: the timed path must really invert (PERIODIC round trip)
    rt = float((xr[0] - xs[0]).abs().max())

    # ---- roofline of the dominant kernel: the fused analysis launch(es), timed alone with CUDA events -------
    peak, peak_src = measured_peaks()
    reps = max(10, args.steps, int(80.0 / max(ms_step / 2.0, 1e-3)))   # >= 80 ms per direction: stable event timing, clock samples
    reps = min(reps, 2000)
    for _ in range(3):
        eng.forward(xs[0], hs, gs, levels, mode, 0, ws[0], vs[0])
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lf0 = eng.launch_count()
    f0.record()
    for i in range(reps):
        k = i % nsets
        eng.forward(xs[k], hs, gs, levels, mode, 0, ws[k], vs[k])
    f1.record()
    torch.cuda.synchronize()
    fwd_ms = f0.elapsed_time(f1) / reps
    fwd_launches = (eng.launch_count() - lf0) // reps
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(reps):
        k = i % nsets
        eng.inverse(ws[k], vs[k], hs, gs, mode, None, 0, out=xr[k])
    g1.record()
    torch.cuda.synchronize()
    inv_ms = g0.elapsed_time(g1) / reps
    # the sampler ran through the timed steps AND the per-direction loops above (the same kernels under load): a 20-step
    # timed region of this workload lasts ~8 ms, too short for nvidia-smi's sampling period on its own
    clocks = sampler.stop() if rank == 0 else None
    alg_bytes_dir = 24.0 * levels * b * n          # 24 B/sample/level/direction
    achieved = alg_bytes_dir / fwd_ms * 1e-6        # GB/s over the forward direction's launches
    roofline = {"bound": "hbm", "kernel": "k_fused_analysis", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_dir / max(fwd_launches, 1),
                "launches_per_direction": fwd_launches, "avg_launch_ms": fwd_ms / max(fwd_launches, 1),
                "forward_ms": fwd_ms, "inverse_ms": inv_ms,
                "inverse_achieved_gbs": alg_bytes_dir / inv_ms * 1e-6,
                "fused_compulsory_bytes_per_direction": 8.0 * (levels + 2) * b * n,
                "frac_of_fused_compulsory_bound": (8.0 * (levels + 2) * b * n / fwd_ms * 1e-6) / peak,
                "fp64_fma_peak_tflops_measured": 37.0,   # tools/dfma_probe.cu on this pool: 63.7 DFMA/clk/SM at 1965 MHz
                "fp64_tflops_achieved": 4.0 * hs.size * levels * b * n / fwd_ms * 1e-9}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(args.workload)
        except Exception:
            pass

    # ---- e2e: host buffers through the public facade, copies inside the timed region (rank-local) ----------
    e2e = None
    if not args.no_e2e:
        eb = b if b * n * 8 * (levels + 2) <= 8e9 else max(1, int(8e9 / (n * 8 * (levels + 2))))
        xh = eng.pinned_empty((eb, n))
        xh[...] = np.random.default_rng(7 + rank).standard_normal((eb, n))
        wh = eng.pinned_empty((levels, eb, n))
        vh = eng.pinned_empty((eb, n))
        oh = eng.pinned_empty((eb, n))

        def e2e_step():
            eng.forward(xh, hs, gs, levels, mode, 0, wh, vh)                 # H2D x, D2H W + V_J
            eng.inverse(wh, vh, hs, gs, mode, None, 0, out=oh)              # H2D W + V_J, D2H x^
        for _ in range(2):
            e2e_step()
        barrier()
        esteps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(esteps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        assert float(np.max(np.abs(oh - xh))) < 1e-6
        e2e = {"value": world * eb * n * esteps / dt * 1e-9, "unit": UNIT,
               "h2d_bytes_per_step": int((levels + 2) * eb * n * 8), "d2h_bytes_per_step": int((levels + 2) * eb * n * 8),
               "api": "Engine.forward/inverse with HOST (pinned) buffers == vw_modwt_forward / vw_modwt_inverse without "
                      "VW_FLAG_DEVICE_PTRS; coefficients cross PCIe both ways like the Java double[] API",
               "batch": eb, "steps": esteps, "numa_node_bound": numa_node}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            cpu = cpu_port(args.workload, len(os.sched_getaffinity(0)) or 1)
            cpu["variants"] = cpu_variants(args.workload)
        except Exception as ex:  # the checker must never take the bench down
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload, "wavelet": wname, "batch_per_gpu": b, "signal_length": n,
                           "levels": levels, "boundary": "PERIODIC", "sharding": "by signal, no communication",
                           "l2": f"{nsets} rotating input/output sets of {(levels + 3) * b * n * 8 / 2**20:.0f} MiB each "
                                 "(> 126 MB L2); inputs larger than L2"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "round_trip_max_abs_err": rt,
                "roofline_model_gsamples": peak / (48.0 * levels),
                "frac_of_roofline_model": value / world / (peak / (48.0 * levels)),
                "gsample_levels_per_s": value * 2.0 * levels}      # 2*J*B*N / t
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
