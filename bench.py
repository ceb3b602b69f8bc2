#!/usr/bin/env python
"""bench.py — headline benchmark of the PDMP ensemble engine (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--traj T]

Workload (BASELINE.json configs[1], "C2"): TCP model (examples/tcp.jl), CHV algorithm, FP64,
10^7 trajectories per GPU, tspan [0, 200], n_jumps cap 1000, reference default tolerances
(reltol 1e-7, abstol 1e-9), post-jump saving of every jump (time, xc, xd), Dormand-Prince 5(4),
synthetic initial conditions xc0_i = 0.5 + u_i, ensemble moments at 4 sample times.

One "step" = one full pass of the hot path over the ensemble:
  value : jumps/s with the initial conditions already resident in HBM; the step is
          [ensemble kernel -> offsets scan -> pool->CSR compaction], results left in HBM.
  e2e   : the same through the host-buffer API: H2D of the step's initial conditions from
          pinned memory, the step, D2H of every saved record + counters + statistics.
N > 1 (torchrun): one process per GPU, trajectories sharded by global id range (weak scaling:
10^7 per GPU), NCCL only for the all-reduce of the statistics buffers.

--impl reference: the reference's CPU path for the same workload.  The reference is Julia and
there is no Julia runtime in this image, so this arm times the CPU oracle port (oracle/,
Tsit5 + the reference's semantics) on all host cores over a bounded sample of trajectories.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "jumps_per_sec_tcp_chv_ensemble"
UNIT = "jumps/s"
TF, N_JUMPS, RELTOL, ABSTOL = 200.0, 1000, 1e-7, 1e-9
SAMPLE_TIMES = [50.0, 100.0, 150.0, 200.0]
FLOPS_PER_ATTEMPT_TCP = 180.0   # D*72 + 2 + 10 + 6*C_rhs, D = 2, C_rhs = 4
REC_BYTES_TCP = 32.0            # 8*(1+n_c) + 8*n_d per saved point; pool record
CSR_BYTES_TCP = 17.0            # CSR / wire record with the event transport: 8 (t) + 8 (xc) + 1 (reaction index)
# DRAM bytes per saved record measured by `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum over the
# records of that launch; profiles/r2_tcp_kernel_final_4M_full.txt (incl. the sample rows), profiles/r2_compact_kernel_final_full.txt)
NCU_DRAM_BYTES_PER_RECORD = {"ensemble_kernel": 34.6, "compact_kernel": 50.9}


def workload_config(n_traj, n_gpus):
    return {
        "workload": "C2: TCP CHV ensemble (examples/tcp.jl), FP64, tspan [0,200], post-jump saving",
        "trajectories_per_gpu": n_traj, "trajectories_total": n_traj * n_gpus,
        "n_jumps_cap": N_JUMPS, "reltol": RELTOL, "abstol": ABSTOL, "ode": "dp5",
        "tstop_policy": "1 on the GPU: after a step clipped by a jump threshold the proposal never drops below the "
                        "pending unclipped one (every step stays error-controlled); the CPU arm keeps OrdinaryDiffEq's rule",
        "save_positions": [False, True], "sample_times": SAMPLE_TIMES,
        "xd_transport": "auto -> event (no Delta and |xd| <= 1 + 999 < 2^23 are provable for TCP with n_jumps = 1000): the wire "
                        "carries the index of the reaction behind every saved point and the binding rebuilds the reference's Int64 "
                        "xd as xd0 + a prefix sum of nu rows on access; a saved point is 8 (t) + 8 (xc) + 1 = 17 B on the wire",
        "initial_conditions": "synthetic xc0_i = 0.5 + u_i (numpy PCG64 seed 2024), xd0 = [0, 1]",
        "rng": "philox4x32-10, seed = step index",
        "l2": "working set (42 GB of pool records + 22 GB of CSR per step) is far larger than the 126 MB L2; no flush needed",
        "parallelism": f"ensemble sharding x{n_gpus} (contiguous global id ranges), stats all-reduce only",
    }


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(device):
    """Pin this rank's host threads (and therefore its first-touch pinned staging) to the CPUs NVML
    reports as local to the GPU, so that 8 ranks streaming results over PCIe do not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def host_ics(n, pinned=True):
    import torch
    u = np.random.default_rng(2024).random(n)
    t = torch.empty((n, 1), dtype=torch.float64, pin_memory=pinned)
    a = t.numpy()
    a[:, 0] = 0.5 + u
    return t, a


def pcie_peak(torch, dist, world, nbytes=1 << 30):
    """Pinned-memory copy peaks of this box, all ranks at once (plain cudaMemcpyAsync through torch): the denominator
    of the end-to-end roofline.  Returns (d2h GB/s aggregate, h2d GB/s aggregate)."""
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = []
    for dst, src in ((h, d), (d, h)):
        best = 0.0
        for rep_ in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); dst.copy_(src, non_blocking=True); b.record()
            torch.cuda.synchronize()
            if rep_:
                best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
        t = torch.tensor([best], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        out.append(float(t.item()))
    del h, d
    return out


def cpu_baseline(n_sample, threads=None, ode="tsit5"):
    """The oracle port on the host cores: a bounded sample of the same workload."""
    import oracle as orc
    import pdmp_b200 as P
    pb = P.models.problem("tcp", tspan=(0.0, TF))
    nthr = threads or orc.max_threads()
    orc.set_threads(nthr)
    xc0 = 0.5 + np.random.default_rng(2024).random((n_sample, 1))
    kw = dict(algorithm="chv", ode=ode, n_jumps=N_JUMPS, reltol=RELTOL, abstol=ABSTOL, xc0=xc0,
              sample_times=SAMPLE_TIMES)
    orc.solve(pb, "tcp", trajectories=min(2000, n_sample), seed=0, **{**kw, "xc0": xc0[:min(2000, n_sample)]})  # warm-up
    t0 = time.perf_counter()
    r = orc.solve(pb, "tcp", trajectories=n_sample, seed=1, **kw)
    dt = time.perf_counter() - t0
    return {"value": r.counters["total_jumps"] / dt, "unit": UNIT, "cores": nthr, "kind": "port",
            "sample": f"{n_sample} trajectories of the same workload ({r.counters['total_jumps']} jumps, {dt:.2f} s), "
                      f"oracle/ C++ restatement with {ode} (the reference is Julia; no Julia runtime in this image)",
            "trajectories_per_sec": n_sample / dt, "seconds": dt}


def run_reference(args):
    """--impl reference: CPU oracle port, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle as orc
    nthr = orc.max_threads()
    cal = cpu_baseline(max(2000, 250 * nthr))
    per_traj = cal["seconds"] / max(2000, 250 * nthr)
    n_sample = int(min(2_000_000, max(4000, 4.0 / per_traj)))   # ~4 s per step
    vals, tps, secs = [], [], []
    for s in range(args.warmup + args.steps):
        b = cpu_baseline(n_sample)
        if s >= args.warmup:
            vals.append(b["value"]); tps.append(b["trajectories_per_sec"]); secs.append(b["seconds"])
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.traj, args.gpus),
        "trajectories_per_sec": float(np.mean(tps)),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": nthr, "kind": "port",
                         "sample": f"{n_sample} trajectories per step of the C2 workload; oracle/ C++ restatement "
                                   "(Tsit5, reference semantics) on all host threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import pdmp_b200 as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if P.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.traj
    pb = P.models.problem("tcp", tspan=(0.0, TF))
    ics_t, ics = host_ics(n)
    def mk(ntr, ic, policy=1):
        return P.Ensemble(pb, P.CHVGPU("tcp", ode="dp5"), trajectories=ntr, traj_id_offset=rank * n,
                          n_jumps=N_JUMPS, reltol=RELTOL, abstol=ABSTOL, save_positions=(False, True),
                          sample_times=SAMPLE_TIMES, device=local, max_records=int(ntr * 160), xc0=ic,
                          tstop_policy=policy, xd_transport="auto")
    # the same resident step under OrdinaryDiffEq's own step rule (tstop_policy 0), reported next to the headline
    value_pol0 = None
    if not args.no_policy0:
        e0 = mk(n, ics, policy=0)
        for s in range(2):
            e0.run(seed=2000 + s)
            e0.counters()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        j0 = at0 = 0
        for s in range(2):
            e0.run(seed=s, stream=torch.cuda.current_stream().cuda_stream)
            c0 = e0.counters(); j0 += c0["total_jumps"]; at0 += c0["rk_attempts"]
        b.record(); torch.cuda.synchronize()
        value_pol0 = {"value": j0 / (a.elapsed_time(b) * 1e-3), "unit": UNIT, "rk_attempts_per_jump": at0 / max(j0, 1), "steps": 2,
                      "note": "rank 0's GPU only; tstop_policy 0 = OrdinaryDiffEq's h_clipped/q after a step clipped by a threshold"}
        e0.close(); del e0
    ens = mk(n, ics)
    n_chunks = 1 if n <= 2 * (1 << 19) else -(-n // (1 << 19))
    stats_t, hist_t = P.distributed.stats_tensors(ens, dev)   # zero-copy torch views of the library's buffers
    stream = torch.cuda.current_stream().cuda_stream

    def step(seed):
        ens.run(seed=seed, stream=stream)
        if world > 1 and stats_t is not None:
            dist.all_reduce(stats_t)      # the only collective on the path: ensemble statistics

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = P.fp64_peak_tflops(local)
    d2h_peak, h2d_peak = pcie_peak(torch, dist, world)
    for s in range(args.warmup):
        step(1000 + s)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    jumps = attempts = records = 0
    kernel_ms = compact_ms = 0.0
    barrier()
    ev0.record()
    for s in range(args.steps):
        step(s)
        c = ens.counters()                # syncs the step; ~tens of microseconds of idle per step
        jumps += c["total_jumps"]; attempts += c["rk_attempts"]; records += c["total_records"]
        kernel_ms += c["kernel_ms"]; compact_ms += c["compact_ms"]
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(jumps), float(n * args.steps)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_max = float(t_ms.item())
    value = float(tot[0].item()) / (ms_max * 1e-3)
    traj_per_s = float(tot[1].item()) / (ms_max * 1e-3)

    # how the trajectories of the last step ended (by tf vs by the n_jumps cap)
    fin = ens.fetch(records=False)
    status_hist = fin.status_histogram()
    mean_jumps = float(fin.njumps.mean() - 1)
    del fin

    # ---- end to end through the host-buffer API (H2D ICs, step, D2H of everything) ----
    # the pinned host staging holds every record of a step (~27 GB at 10^7 trajectories): bound it
    # by the host memory actually available to this rank
    stats_t = hist_t = None
    ens.close()
    import psutil
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    budget = 0.5 * psutil.virtual_memory().available / max(local_world, 1)
    n_e2e = int(min(n, budget / (160 * CSR_BYTES_TCP * 1.02)))
    n_e2e = max(1 << 16, (n_e2e >> 16) << 16) if n_e2e < n else n
    ens = mk(n_e2e, ics[:n_e2e])
    ics_e = ics[:n_e2e]
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    r = ens.run_host(seed=999, xc0=ics_e)   # warm-up: allocates (pins) the host staging
    del r
    barrier()
    t0 = time.perf_counter()
    ej = 0
    for s in range(e2e_steps):
        r = ens.run_host(seed=s, xc0=ics_e)   # H2D ICs -> chunked kernels -> D2H of every record, overlapped
        if world > 1:
            P.distributed.allreduce_stats_host(r)   # global statistics on every rank
        ej += r.counters["total_jumps"]
        xd_wire = r._xd_raw if r._event is not None else r.xd     # (touching r.xd would rebuild it from the event bytes)
        d2h_bytes = (r.time.nbytes + r.xc.nbytes + xd_wire.nbytes + r.offsets.nbytes + r.njumps.nbytes +
                     r.nrejected.nbytes + r.status.nbytes + r.batch_count.nbytes * 3)
        del r
    barrier()
    e2e_s = time.perf_counter() - t0
    e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    ejt = torch.tensor([float(ej)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        dist.all_reduce(ejt)
    e2e_value = float(ejt.item()) / float(e.item())

    if rank == 0:
        ach = attempts * FLOPS_PER_ATTEMPT_TCP / (kernel_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        comp_gbs = records * (REC_BYTES_TCP + CSR_BYTES_TCP) / (compact_ms * 1e-3) / 1e9 if compact_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(n, world),
            "trajectories_per_sec": traj_per_s,
            "jumps_per_step_rank0": jumps / args.steps, "records_per_step_rank0": records / args.steps,
            "rk_attempts_per_jump": attempts / max(jumps, 1),
            "status_rank0_last_step": status_hist, "mean_jumps_per_trajectory": mean_jumps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ics_e.nbytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps, "trajectories_per_gpu": n_e2e,
                    "ms_per_step": 1e3 * float(e.item()) / e2e_steps},
            "roofline_e2e": {"bound": "pcie", "achieved": world * d2h_bytes / (float(e.item()) / e2e_steps) / 1e9, "peak": d2h_peak,
                             "unit": "GB/s", "frac": world * d2h_bytes / (float(e.item()) / e2e_steps) / 1e9 / d2h_peak,
                             "h2d_peak": h2d_peak,
                             "note": "device->host bytes of one e2e step (all ranks) / its wall time, against the pinned D2H copy "
                                     "peak of this box measured in this run with all ranks copying at once"},
            "value_tstop_policy0": value_pol0,
            "gpu_launches": 4 * args.steps,
            "gpu_launches_note": "timed (resident) region, per step: ensemble_kernel + stats_reduce_kernel + chunk_close_kernel + "
                                 "compact_kernel, ours; plus cub::DeviceScan and memsets.  The e2e region launches the same four per "
                                 f"chunk ({n_chunks} chunks per step)",
            "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                         "traffic": NCU_DRAM_BYTES_PER_RECORD["ensemble_kernel"] * records / args.steps,
                         "traffic_note": "bytes per launch = ncu-measured DRAM bytes per saved record (profiles/) x records of one "
                                         "launch; algorithmic output bytes are 32 B/record, the integrator state never leaves registers",
                         "kernel": "ensemble_kernel<TcpDev,TabDP5,CHV,records,stats>",
                         "peak_source": "live DFMA microbenchmark in libpdmp_b200.so (MEASURED_PEAKS.json has no FP64 entry)",
                         "algorithmic_flops_per_rk_attempt": FLOPS_PER_ATTEMPT_TCP,
                         "rk_attempts_per_step": attempts / args.steps, "kernel_ms_per_step": kernel_ms / args.steps},
            "roofline_output": {"bound": "hbm", "kernel": "compact_kernel<TcpDev> (pool -> CSR)", "achieved": comp_gbs,
                                "peak": hbm_peak, "unit": "GB/s", "frac": (comp_gbs / hbm_peak) if comp_gbs else None,
                                "algorithmic_bytes_per_record": REC_BYTES_TCP + CSR_BYTES_TCP,
                                "traffic": NCU_DRAM_BYTES_PER_RECORD["compact_kernel"] * records / args.steps,
                                "compact_ms_per_step": compact_ms / args.steps,
                                "note": "compact_ms also contains the offsets scan"},
            "clocks": clk,
        }
        if numa:
            line["host_affinity_rank0"] = f"{len(numa)} CPUs local to GPU {local} (NVML)"
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(args.cpu_sample).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    ens.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_single_process(args):
    """--single-process: ONE process drives all N GPUs through the library's own sharding (pdmp_solve_opts.devices):
    one host thread per device inside libpdmp_b200.so, statistics summed there.  The same workload and step as the
    torchrun path, for comparison with it (SURVEY.md 8(b) "device list")."""
    import torch
    import pdmp_b200 as P
    G = args.gpus
    if P.device_count() < G:
        raise RuntimeError(f"--single-process --gpus {G}: only {P.device_count()} CUDA devices")
    n = args.traj * G
    pb = P.models.problem("tcp", tspan=(0.0, TF))
    ics_t, ics = host_ics(n)
    ens = P.Ensemble(pb, P.CHVGPU("tcp", ode="dp5"), trajectories=n, n_jumps=N_JUMPS, reltol=RELTOL, abstol=ABSTOL,
                     save_positions=(False, True), sample_times=SAMPLE_TIMES, devices=list(range(G)),
                     max_records=int(n * 160), xc0=ics, tstop_policy=1, xd_transport="auto")
    sync = lambda: [torch.cuda.synchronize(g) for g in range(G)]
    for s in range(args.warmup):
        ens.run(seed=1000 + s); ens.counters()
    sync()
    clocks = ClockSampler(0); clocks.start()
    t0 = time.perf_counter()
    jumps = attempts = 0
    kernel_ms = 0.0
    for s in range(args.steps):
        ens.run(seed=s)
        c = ens.counters()        # waits for every shard: device work of the step is complete
        jumps += c["total_jumps"]; attempts += c["rk_attempts"]; kernel_ms += c["kernel_ms"]
    sync()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    r = ens.run_host(seed=999, xc0=ics); del r
    t1 = time.perf_counter(); ej = 0
    for s in range(e2e_steps):
        r = ens.run_host(seed=s, xc0=ics); ej += r.counters["total_jumps"]; del r
    e2e_s = time.perf_counter() - t1
    line = {"metric": METRIC, "value": jumps / dt, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(workload_config(args.traj, G), mode="single process, opts.devices (one host thread "
                                                                                   "per GPU inside the library)"),
            "trajectories_per_sec": n * args.steps / dt, "rk_attempts_per_jump": attempts / max(jumps, 1),
            "kernel_ms_per_step_max_over_devices": kernel_ms / args.steps,
            "e2e": {"value": ej / e2e_s, "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": int(ics.nbytes)},
            "gpu_launches": 4 * args.steps * G, "clocks": clk,
            "timing": "host wall clock around K steps, every device synchronised on both sides"}
    print(json.dumps(line))
    ens.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--traj", type=int, default=10_000_000, help="trajectories per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=1_500_000, help="trajectories of the CPU baseline sample (~15 s on 16 cores)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-policy0", action="store_true", help="skip the extra resident measurement under tstop_policy 0")
    ap.add_argument("--single-process", action="store_true",
                    help="N > 1 without torchrun: ONE process, the ensemble sharded over N GPUs inside the library call (opts.devices)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.single_process and args.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return run_single_process(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
