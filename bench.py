#!/usr/bin/env python
"""bench.py -- FD forward+adjoint throughput of the B200 path (and of the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    torchrun --nproc-per-node N ... bench.py --gpus N ...          (one rank per GPU, NCCL)

One "step" = one gradient evaluation of the hot path over one batch of synthetic velocity models: forward modelling of
every shot (nt levels) + the reverse-time adjoint with the imaging condition, i.e. B*ns*nzp*nxp*nt forward+adjoint
cell-update *pairs* (SURVEY.md 8d).

Prints ONE JSON line.  Top level = the headline: BASELINE.json configs[1] (OpenFWI 70x70, 64 models x 5 shots x 1000
levels) per GPU -- models are independent, so N GPUs are N x 64 models with no data-path collective ("scaling": "weak").
`value` is measured with inputs resident in HBM, `e2e` through the public operator (FWIForward + autograd) from pinned
host buffers.  Blocks beside the headline, every one measured in this run:

  "sharded"       the north-star's multi-GPU path: ONE Marmousi-shaped model x 176 shots dealt over the N ranks by
                  ShardedFWIForward, the gradient all-reduce inside the timed region (STRONG scaling: the same work at every
                  N, so ms_per_step at N = 1, 2, 4, 8 gives the scaling efficiency); parity of the sharded gradient against
                  the reference fixture (the reference's own 5 shots dealt 3+2 / 2+1+1+1 / 1x5+idle over the live ranks)
  "red_iter"      (N = 1) seconds per full RED-DiffEq iteration -- FD solve + the reference's own U-Net regulariser
                  (staged unmodified under baseline/_ref, random-init) + Adam + metrics -- for 64 OpenFWI models and for
                  one Marmousi model (3 patches), with the reference's regulariser call pattern and with this repo's
  "cpu_baseline"  the oracle C port (OpenMP, every host core) on a bounded sample of the headline workload
  "reference_pde" the UNMODIFIED reference operator (baseline/_ref/red_diffeq/solvers/pde.py) on the host CPU, one model
"""
import argparse
import csv
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO_BYTES_FWD = 12.0   # read p_{t-1}, p_{t-2}; write p_t                      
ALGO_BYTES_ADJ = 16.0   # read q_{t+1}, q_{t+2}, p_{t-1}; write q_t
ALGO_BYTES_PAIR = ALGO_BYTES_FWD + ALGO_BYTES_ADJ

WORKLOADS = {
    # name: (pde ctx factory, nz, nx, models per GPU)
    "openfwi_b64": ("openfwi", 70, 70, 64),       # BASELINE.json configs[1]: the headline (weak scaling: 64 models per GPU)
    "openfwi_b1": ("openfwi", 70, 70, 1),         # configs[0]
    "marmousi_b1": ("marmousi", 70, 190, 1),      # configs[2] shape, reference shot count (5)
    "marmousi_b16": ("marmousi", 70, 190, 16),
    # configs[2]: ONE Marmousi-shaped model, 176 shots (the reference's 5 do not divide over 8 GPUs; 176 =
    # 8 GPUs x one wave of 22 co-resident six-CTA clusters, what cudaOccupancyMaxActiveClusters reports on a B200)
    # sharded over the ranks by ShardedFWIForward: strong scaling, one gradient all-reduce per step
    "marmousi_sharded": ("marmousi176", 70, 190, 1),
    # configs[3]: Overthrust shape (= the Marmousi grid in the reference's configs) with a long record, nt = 4000
    # (synthetic extension); run with --history-segment K to exercise wavefield checkpointing
    "overthrust_long": ("overthrust4000", 70, 190, 8),
    "overthrust_16000": ("overthrust16000", 70, 190, 2),
}

# ncu launch lists of the bench command (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
# --clock-control none --csv ... python bench.py --only-headline --workload W`), committed under profiles/: the source of
# `roofline.traffic` (DRAM bytes of ONE launch of a kernel class).  Not measured in the bench run itself -- the key says so.
TRAFFIC_PROFILES = {
    "openfwi_b64": "profiles/launches_r2b_openfwi_b64.csv",
    "marmousi_b1": "profiles/launches_r2b_marmousi_b1.csv",
    "marmousi_sharded": "profiles/launches_r2b_marmousi_sharded.csv",
    "overthrust_long": "profiles/launches_r2b_overthrust_long.csv",
}


def make_ctx(kind):
    from red_diffeq_b200.utils import synthetic
    if kind == "openfwi":
        return dict(synthetic.PDE_OPENFWI)
    ctx = dict(synthetic.PDE_MARMOUSI)
    if kind == "marmousi176":
        ctx["ns"] = 176
    if kind == "overthrust4000":
        ctx["nt"] = 4000
    if kind == "overthrust16000":
        ctx["nt"] = 16000
    return ctx


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def classify_kernel(name):
    """Kernel class of a demangled kernel name as ncu prints it (None = not one of the time-loop kernels)."""
    if "k_fwd_cluster<" in name:
        targs = name.split("k_fwd_cluster<", 1)[1].split(">", 1)[0].replace(" ", "").split(",")
        mode = targs[2].replace("(int)", "").replace("(bool)", "")
        return "adjoint_resident" if mode == "2" else ("adjoint_field" if mode in ("true", "1") else "forward")
    if "k_step_tile<" in name:
        targs = name.split("k_step_tile<", 1)[1].split(">", 1)[0].replace(" ", "").split(",")
        return "adjoint_field" if targs[-1] in ("true", "(bool)1", "1") else "forward"
    if "k_imaging" in name:
        return "imaging"
    if "k_adj_step" in name:
        return "adjoint_loop"
    return None


def load_traffic_profile(path):
    """Per kernel class: DRAM bytes (read + write) of one launch = the mean over the launches of that class in an ncu
    --csv launch list (the launches of a class are identical except in the recompute tier, whose modelling forward writes
    no history and whose recomputing forward does: the mean x launches is then still the class's bytes per step), plus the
    grid / block sizes ncu recorded."""
    full = os.path.join(ROOT, path)
    with open(full, newline="") as f:
        rows = list(csv.reader(f))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    col = {name: i for i, name in enumerate(rows[hi])}
    launches = {}
    for r in rows[hi + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        e = launches.setdefault(int(r[col["ID"]]), {"name": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
        e[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
    out = {}
    for e in launches.values():
        cls = classify_kernel(e["name"])
        if cls is None or "dram__bytes_read.sum" not in e:
            continue
        c = out.setdefault(cls, {"bytes": [], "ns": [], "grid": set(), "block": set(), "kernel": e["name"].split("(")[0][-60:]})
        c["bytes"].append(e["dram__bytes_read.sum"] + e.get("dram__bytes_write.sum", 0.0))
        c["ns"].append(e.get("gpu__time_duration.sum", 0.0))
        c["grid"].add(e["grid"])
        c["block"].add(e["block"])
    return {k: {"bytes_per_launch": float(np.mean(v["bytes"])), "ncu_us_per_launch": float(np.mean(v["ns"])) * 1e-3,
                "launches_in_profile": len(v["bytes"]), "grid": sorted(v["grid"]), "block": sorted(v["block"]), "kernel": v["kernel"]}
            for k, v in out.items()}


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm_all, sm_busy, smax, reasons = [], [], None, set()
        for r in self.rows:
            try:
                clk = float(r[0]); smax = float(r[1])
            except Exception:
                continue
            sm_all.append(clk)
            try:   # "under load" = the samples nvidia-smi itself reports a busy GPU for (the sampler spans the timed regions only)
                if float(r[7]) >= 50.0:
                    sm_busy.append(clk)
            except Exception:
                pass
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        use = sm_busy if sm_busy else sm_all
        return {"sm_mhz": float(np.median(use)) if use else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm_all), "samples_under_load": len(sm_busy)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def cpu_baseline(ctx, nz, nx, target_seconds=20.0):
    """Times the oracle port (CPU restatement of the reference algorithm) on a bounded sample of the workload."""
    from oracle import build_oracle, fwi_oracle
    from red_diffeq_b200.utils import synthetic
    build_oracle.build()
    # all the host cores this process may use: launchers such as torchrun export OMP_NUM_THREADS=1 to their workers, which
    # would time the CPU arm on one thread at N > 1
    fwi_oracle.set_threads(host_threads())
    threads = fwi_oracle.threads()
    sv = fwi_oracle.Survey(dict(ctx), nz, nx)
    B = max(1, -(-threads // sv.ns))          # one shot per host thread
    B = min(B, 32)
    vn = synthetic.velocity_models(B, nz, nx, seed=99)
    v_phys = (vn + np.float32(1)) / np.float32(2) * np.float32(3000) + np.float32(1500)
    cot = synthetic.cotangent((B, sv.ns, sv.nt_out, sv.nrec), seed=100)
    best, reps, t_total = None, 0, 0.0
    while reps < 3 and (reps == 0 or t_total < target_seconds):
        t0 = time.perf_counter()
        fwi_oracle.gradient(sv, v_phys, cot)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        reps += 1
        t_total += dt
    pairs = sv.pairs(B)
    return {"value": pairs / best, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": f"oracle C port (OpenMP, {threads} threads), {B} models x {sv.ns} shots x {sv.nt} levels, "
                      f"fwd+adjoint, best of {reps} ({best:.2f} s)"}, pairs, best


def reference_pde_leg(kind="openfwi", nz=70, nx=70, reps=3):
    """The UNMODIFIED reference operator (red_diffeq/solvers/pde.py:6-93, staged under baseline/_ref by
    baseline/stage_reference.py) on the host CPU: one model, fp32, forward + autograd backward with a fixed cotangent, one
    warm-up call then best of `reps`.  Its autograd tape is ~12 GB per OpenFWI model, so one model is the
    sample; throughput is per pair, comparable with every other number here."""
    try:
        from baseline import ref_loader
        if not ref_loader.available():
            return {"unavailable": "baseline/_ref is empty (run baseline/stage_reference.py where /root/reference exists)"}
        import torch
        from red_diffeq_b200.utils import synthetic
        from red_diffeq_b200.utils.data_trans import s_normalize_none, v_denormalize
        pde = ref_loader.load("red_diffeq.solvers.pde")
        threads = host_threads()
        torch.set_num_threads(threads)
        ctx = make_ctx(kind)
        op = pde.FWIForward(dict(ctx), torch.device("cpu"), normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
        vn = torch.from_numpy(synthetic.velocity_models(1, nz, nx, seed=99))
        cot = torch.from_numpy(synthetic.cotangent((1, ctx["ns"], ctx["nt"], ctx["ng"]), seed=100))
        fwd, bwd = [], []
        for i in range(reps + 1):
            v = vn.clone().requires_grad_(True)
            t0 = time.perf_counter()
            seis = op(v)
            t1 = time.perf_counter()
            (seis * cot).sum().backward()
            t2 = time.perf_counter()
            if i > 0:
                fwd.append(t1 - t0); bwd.append(t2 - t1)
            del seis, v
        pairs = float(ctx["ns"]) * (nz + 2 * ctx["nbc"]) * (nx + 2 * ctx["nbc"]) * ctx["nt"]
        best = min(f + b for f, b in zip(fwd, bwd))
        cpu = ""
        try:
            with open("/proc/cpuinfo") as f:
                cpu = next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), "")
        except OSError:
            pass
        return {"value": pairs / best, "unit": "pairs/s", "kind": "reference", "cores": threads, "torch_threads": torch.get_num_threads(),
                "cpu": cpu, "forward_s": min(fwd), "backward_s": min(bwd), "best_s": best,
                "sample": f"unmodified red_diffeq/solvers/pde.py FWIForward on CPU, 1 {kind} model x {ctx['ns']} shots x {ctx['nt']} levels, "
                          f"fp32, forward + autograd backward, 1 warm-up + best of {reps}"}
    except Exception as e:   # a reported baseline must never take the headline down with it
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference algorithm on the host CPU.  `value` = the oracle C port with OpenMP on every host core
    (the FASTER of the two CPU implementations: a conservative denominator); `reference_pde` = the unmodified Python
    reference timed beside it on one model."""
    if rank != 0:
        return
    kind, nz, nx, B = WORKLOADS[args.workload]
    ctx = make_ctx(kind)
    times, pairs, base = [], None, None
    for i in range(args.warmup + args.steps):
        base, pairs, best = cpu_baseline(ctx, nz, nx, target_seconds=0.0)
        if i >= args.warmup:
            times.append(best)
    t = float(np.mean(times))
    value = pairs / t
    base["value"] = value
    line = {"impl": "reference", "metric": "FD cell-updates/s (fwd+adjoint)", "value": value, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "note": "bounded CPU sample of the same workload: " + base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_reference_pde:
        line["reference_pde"] = reference_pde_leg(kind if kind in ("openfwi", "marmousi") else "marmousi", nz, nx)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self):
        import torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, values):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return list(values)
        t = torch.tensor(list(values), device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(self, value):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return float(value)
        t = torch.tensor([float(value)], device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()


def measure_workload(env, workload, steps, warmup, opts=(), history_segment=None, batch=0, nt=0, ns=0, sampler=None):
    """Times `steps` gradient evaluations of `workload` (resident inputs, then end to end from pinned host buffers); returns
    the measurements of this rank with the cross-rank maxima / sums already applied."""
    import torch
    import torch.distributed as dist
    from red_diffeq_b200 import FWIForward, ShardedFWIForward, s_normalize_none, v_denormalize
    from red_diffeq_b200.utils import synthetic
    dev, rank, world = env.dev, env.rank, env.world
    kind, nz, nx, B = WORKLOADS[workload]
    if batch:
        B = batch
    ctx = make_ctx(kind)
    if nt:
        ctx["nt"] = nt
    if ns:
        ctx["ns"] = ns
    sharded = workload == "marmousi_sharded"
    ns, nt, nbc = ctx["ns"], ctx["nt"], ctx["nbc"]
    nzp, nxp = nz + 2 * nbc, nx + 2 * nbc
    if sharded:
        # strong scaling: the same 1 x 176 shots whatever the rank count; every rank models its shots
        wrapper = ShardedFWIForward(dict(ctx), dev, mode="shots", normalize=True, v_denorm_func=v_denormalize,
                                    s_norm_func=s_normalize_none)
        _, _, my_shots = wrapper.partition(B)
        op = wrapper._operator(my_shots)
        ns_local = len(my_shots)
    else:
        wrapper = None
        op = FWIForward(dict(ctx), dev, normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
        ns_local = ns
    for kv in opts:
        k, v = kv.split("=")
        op.set_option(k, int(v))
    if history_segment is not None:
        op.set_history_segment(history_segment)
    pairs_rank = B * ns_local * nzp * nxp * nt
    cells_level = B * ns_local * nzp * nxp

    # synthetic inputs (seed 8888 + rank): models, and "observed" data = a fixed random record so that the
    # L1 misfit of the e2e path has a non-trivial cotangent
    model_seed = synthetic.SEED if sharded else synthetic.SEED + rank   # sharded: the model is replicated
    vn_host = torch.from_numpy(synthetic.velocity_models(B, nz, nx, seed=model_seed)).pin_memory()
    y_host = torch.from_numpy(synthetic.cotangent((B, ns_local, nt, ctx["ng"]), seed=17 + rank)).pin_memory()
    grad_host = torch.empty((B, 1, nz, nx), dtype=torch.float32).pin_memory()
    loss_host = torch.empty((B,), dtype=torch.float32).pin_memory()
    v_dev = vn_host.to(dev)
    cot_dev = y_host.to(dev)
    fwd = wrapper if sharded else op   # the sharded wrapper adds the gradient all-reduce to backward()
    plan = op._plan_for(nz, nx, dev)

    def step_resident():
        v = v_dev.detach().requires_grad_(True)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        seis = fwd(v)
        launches_f = op.last_launches
        e1.record()
        seis.backward(cot_dev)
        e2.record()
        return e0, e1, e2, launches_f, op.last_launches - launches_f

    copy_stream = torch.cuda.Stream(device=dev)

    def step_e2e():
        # the observed data (90 MB for the headline batch) are not needed before the misfit: their host -> device copy runs
        # on a second stream, inside the timed region, while the forward kernel is busy; the velocity batch goes first
        v = vn_host.to(dev, non_blocking=True).requires_grad_(True)
        copy_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            y = y_host.to(dev, non_blocking=True)
        seis = fwd(v)
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
        y.record_stream(torch.cuda.current_stream(dev))
        loss = (seis - y).abs().mean(dim=(1, 2, 3))          # the reference's L1 data misfit (core/losses.py:27-41)
        loss.sum().backward()
        grad_host.copy_(v.grad, non_blocking=True)
        loss_host.copy_(loss.detach(), non_blocking=True)

    # ---- resident-input timing (value) -------------------------------------------------------------
    for _ in range(warmup):
        step_resident()
    env.barrier()
    plan.set("timing", 1)   # library-side CUDA events per kernel class, inside the timed region
    if sampler is not None:
        sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_stop = torch.cuda.Event(enable_timing=True)
    t_start.record()
    evs = [step_resident() for _ in range(steps)]
    t_stop.record()
    env.barrier()
    elapsed_ms = t_start.elapsed_time(t_stop)
    classes = ("forward", "adjoint_field", "imaging", "adjoint_loop", "adjoint_resident")
    kernel_us = {k: plan.get("us_" + k) for k in classes}
    kernel_n = {k: plan.get("n_" + k) for k in classes}
    plan.set("timing", 0)
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    adj_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    launches_f, launches_b = evs[-1][3], evs[-1][4]

    # ---- end-to-end timing from pinned host buffers (e2e) -------------------------------------------
    for _ in range(max(1, warmup)):   # (as many warm-up steps as the resident-input timing: the allocator pools of the e2e path)
        step_e2e()
    env.barrier()
    e_start = torch.cuda.Event(enable_timing=True)
    e_stop = torch.cuda.Event(enable_timing=True)
    e_start.record()
    for _ in range(steps):
        step_e2e()
    e_stop.record()
    env.barrier()
    e2e_ms = e_start.elapsed_time(e_stop)
    clocks = sampler.stop() if sampler is not None else None

    # the gradient all-reduce alone (sharded workload): what one NCCL all-reduce of this size costs on this box
    allreduce_us = 0.0
    if sharded and world > 1:
        g = torch.zeros((B, 1, nz, nx), device=dev)
        for _ in range(5):
            dist.all_reduce(g)
        env.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            dist.all_reduce(g)
        a1.record()
        env.barrier()
        allreduce_us = a0.elapsed_time(a1) * 1e3 / 20

    elapsed_ms, e2e_ms, allreduce_us = env.max_over_ranks([elapsed_ms, e2e_ms, allreduce_us])
    pairs_total = env.sum_over_ranks(pairs_rank)
    seg = plan.get("history_segment")
    res = {
        "workload": workload, "B": B, "ns": ns, "ns_local": ns_local, "nt": nt, "nz": nz, "nx": nx, "nzp": nzp, "nxp": nxp,
        "pairs_rank": pairs_rank, "pairs_total": pairs_total, "cells_level": cells_level,
        "ms_per_step": elapsed_ms / steps, "e2e_ms_per_step": e2e_ms / steps, "fwd_ms": fwd_ms, "adj_ms": adj_ms,
        "kernel_us": {k: kernel_us[k] / steps for k in classes}, "kernel_n": {k: kernel_n[k] // steps for k in classes},
        "launches_f": launches_f, "launches_b": launches_b, "clocks": clocks, "allreduce_us": allreduce_us,
        "h2d": int(vn_host.numel() * 4 + y_host.numel() * 4), "d2h": int(grad_host.numel() * 4 + loss_host.numel() * 4),
        "options": dict(op.options), "segment": seg,
        "plan": {k: plan.get(k) for k in ("adj_split", "cluster_size_used", "cluster_size_last",
                                           "cluster_rows_last", "u_chunk_used", "cluster_wave")},
        "engine_opt": op.options.get("engine", 0),
        "resident_gb": (plan.history_bytes(B, seg) + plan.workspace_bytes(B)) / 1e9,
    }
    op.release_memory()
    del v_dev, cot_dev, fwd, wrapper, op, plan
    torch.cuda.empty_cache()
    return res


def roofline_block(m, steps):
    """Per-kernel-class device times (CUDA events the library records on its launch streams inside the timed region,
    rdfwi_plan_set "timing") against algorithmic bytes, and -- from the committed ncu launch list of the same command -- the
    bytes that really crossed the HBM pins."""
    peak, peak_src = measured_peak_gbs()
    p = m["plan"]
    nt = m["nt"]
    recompute = p["adj_split"] in (2, 5)  # no history kept: the backward pass re-runs the forward kernel per chunk
    resident = p["adj_split"] in (4, 5)   # cluster engine, imaging sums formed inside the adjoint sweep (tensor memory)
    fwd_cluster = m["engine_opt"] != 1 and p["cluster_size_used"] > 0 and (m["segment"] == 0 or recompute)
    adj_split = p["adj_split"] >= 1
    us, n = m["kernel_us"], m["kernel_n"]
    cell_updates = float(m["cells_level"]) * nt
    kernels = {
        "forward": {"kernel": "k_fwd_cluster<EXACT>" if fwd_cluster else "k_step_tile<EXACT>", "us": us["forward"],
                    "launches": n["forward"] if fwd_cluster else m["launches_f"] - 3,
                    # recompute tier: the forward kernel really runs twice per step (modelling + per-chunk recompute)
                    "algo_bytes": ALGO_BYTES_FWD * cell_updates * (2 if recompute else 1)},
    }
    if resident:
        # one kernel: adjoint field on chip, forward history read once, imaging accumulators in tensor memory
        kernels["adjoint_resident"] = {"kernel": "k_fwd_cluster<ADJ+IMAGING>", "us": us["adjoint_resident"], "launches": n["adjoint_resident"],
                                       "algo_bytes": ALGO_BYTES_ADJ * cell_updates}
    elif adj_split:
        # the adjoint's 16 B / cell-update split as: adjoint-field kernel (read u_{t+1}, u_{t+2}, write u_t = 12 B; all of
        # it stays in shared memory, only the 4 B history write reaches HBM) + imaging kernel (pointwise: read p_t and
        # u_t once each = 8 B, which is exactly what it streams from HBM)
        tiled = p["adj_split"] == 3   # per-level engine: nt tiled launches per chunk of shots
        kernels["adjoint_field"] = {"kernel": "k_step_tile<ADJ>" if tiled else "k_fwd_cluster<ADJ>", "us": us["adjoint_field"],
                                    "launches": n["adjoint_field"] * (nt if tiled else 1), "algo_bytes": 12.0 * cell_updates}
        kernels["imaging"] = {"kernel": "k_imaging", "us": us["imaging"], "launches": n["imaging"], "algo_bytes": 8.0 * cell_updates}
    else:
        kernels["adjoint_loop"] = {"kernel": "k_adj_step", "us": us["adjoint_loop"], "launches": m["launches_b"] - 7,
                                   "algo_bytes": ALGO_BYTES_ADJ * cell_updates}
    # DRAM traffic per launch from the committed ncu launch list of this workload
    prof_path, prof, prof_err = TRAFFIC_PROFILES.get(m["workload"]), None, None
    if prof_path is None:
        prof_err = "no ncu launch list committed for this workload"
    else:
        try:
            prof = load_traffic_profile(prof_path)
        except Exception as e:
            prof_err = f"{prof_path}: {type(e).__name__}: {e}"
    for name, kk in kernels.items():
        kk["achieved_gbs"] = kk["algo_bytes"] / (kk["us"] * 1e-6) / 1e9 if kk["us"] > 0 else None
        kk["avg_launch_us"] = kk["us"] / max(kk["launches"], 1)
        kk["frac"] = kk["achieved_gbs"] / peak if kk["achieved_gbs"] else None
        kk["traffic"], kk["ncu_gbs"] = None, None
        if prof is not None:
            c = prof.get(name)
            # the launch list must describe the kernel this run launched: same class present, and its duration per launch
            # within a factor 2 of what this run measured (ncu serialises and runs cold; the SHARE must agree, not the value)
            if c is None:
                prof_err = f"{prof_path} has no launches of class {name}: stale profile"
            elif kk["avg_launch_us"] > 0 and not (0.5 <= c["ncu_us_per_launch"] / kk["avg_launch_us"] <= 2.0):
                prof_err = (f"{prof_path}: {name} took {c['ncu_us_per_launch']:.0f} us per launch under ncu, {kk['avg_launch_us']:.0f} us here: "
                            "the profile is of another launch shape")
            else:
                kk["traffic"] = c["bytes_per_launch"]
                kk["ncu_gbs"] = c["bytes_per_launch"] * kk["launches"] / (kk["us"] * 1e-6) / 1e9 if kk["us"] > 0 else None
    if prof_err is not None:
        print(f"bench.py: roofline.traffic unavailable: {prof_err}", file=sys.stderr, flush=True)
        for kk in kernels.values():
            kk["traffic"], kk["ncu_gbs"] = None, None
    dom_key = max(kernels, key=lambda q: kernels[q]["us"])
    dom = kernels[dom_key]
    ms_per_step = m["ms_per_step"]
    fwd_achieved = ALGO_BYTES_FWD * cell_updates / (m["fwd_ms"] * 1e-3) / 1e9
    adj_achieved = ALGO_BYTES_ADJ * cell_updates / (m["adj_ms"] * 1e-3) / 1e9
    step_bytes = None
    if all(kk["traffic"] is not None for kk in kernels.values()):
        step_bytes = sum(kk["traffic"] * kk["launches"] for kk in kernels.values())
    return {
        "bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
        "frac": dom["frac"], "frac_of_nominal_8TBs": dom["achieved_gbs"] / 8000.0 if dom["achieved_gbs"] else None,
        "traffic": dom["traffic"],
        "traffic_source": (prof_path + " (ncu launch list of this command, committed; per launch, not measured in this run)") if prof_err is None else None,
        "traffic_error": prof_err,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": dom["algo_bytes"] / max(dom["launches"], 1),
        "avg_launch_us": dom["avg_launch_us"], "launches_per_step": dom["launches"],
        "share_of_step": dom["us"] * 1e-3 / ms_per_step,
        "kernels": {k: {"kernel": v["kernel"], "ms_per_step": v["us"] * 1e-3, "launches_per_step": v["launches"],
                        "algorithmic_GB_per_step": v["algo_bytes"] / 1e9, "achieved_gbs": v["achieved_gbs"], "frac": v["frac"],
                        "traffic": v["traffic"], "ncu_dram_gbs_from_profile": v["ncu_gbs"],
                        "ncu_dram_frac_from_profile": v["ncu_gbs"] / peak if v["ncu_gbs"] else None}
                    for k, v in kernels.items()},
        # SURVEY.md 8(d) aggregates: 12 B forward, 16 B adjoint (both adjoint kernels together), 28 B pair
        "forward_frac": fwd_achieved / peak, "adjoint_frac": adj_achieved / peak,
        "pair_frac": (ALGO_BYTES_PAIR * m["pairs_rank"] / ((m["fwd_ms"] + m["adj_ms"]) * 1e-3) / 1e9) / peak,
        # what really crossed the HBM pins over the whole step (profile bytes x launches) / step time / peak: the figure that
        # says how far the step is from the HBM roof of THIS design (the cluster kernels keep their fields on chip)
        "real_dram_GB_per_step": step_bytes / 1e9 if step_bytes else None,
        "frac_real_dram": (step_bytes / (ms_per_step * 1e-3) / 1e9) / peak if step_bytes else None,
    }


def config_block(m):
    p = m["plan"]
    recompute = p["adj_split"] in (2, 5)
    fwd_cluster = m["engine_opt"] != 1 and p["cluster_size_used"] > 0 and (m["segment"] == 0 or recompute)
    adj_split = p["adj_split"] >= 1
    seg = m["segment"]
    adj_txt = "per-level fused"
    if p["adj_split"] in (4, 5):
        adj_txt = "cluster-resident adjoint field with the imaging sums formed in the sweep (accumulators in tensor memory; C=%d)" % p["cluster_size_last"]
    elif adj_split:
        adj_txt = ("split: per-level tiled adjoint field + streaming imaging" if p["adj_split"] == 3 else
                   "split: cluster-resident adjoint field (C=%d) + streaming imaging" % p["cluster_size_last"])
    return {"workload": m["workload"], "models_per_gpu": m["B"], "shots_per_model": m["ns"], "shots_on_rank0": m["ns_local"],
            "nt": m["nt"], "padded_grid": [m["nzp"], m["nxp"]], "pairs_per_step_per_gpu": m["pairs_rank"],
            "l2_policy": "working set (wavefield histories, %.1f GB streamed per step) far exceeds the 126 MB L2; no flush needed" % m["resident_gb"],
            "history": ("none kept: forward field recomputed per chunk of %d shots in the backward pass" % p["u_chunk_used"]) if recompute
                       else (("checkpoint pairs every %d levels" % seg) if seg else "every level"),
            "engine": {"forward": "cluster-resident (C=%d, %d rows per thread)" % (p["cluster_size_last"], p["cluster_rows_last"]) if fwd_cluster else "per-level",
                       "adjoint": adj_txt},
            "options": m["options"]}


# ---------------------------------------------------------------------------------------------------------------------
def sharded_parity(env):
    """The reference's own Marmousi survey (ONE model, ns = 5: configs/marmousi/red-diffeq.yaml:5-15) dealt over the live
    ranks by ShardedFWIForward (3+2 at N = 2, 2+1+1+1 at N = 4, 1x5 + 3 idle ranks at N = 8) against the fixture the
    reference's own pde.py wrote (tests/golden/marmousi.npz: autograd gradient, fixed cotangent): relative L2 of the
    all-reduced gradient (tolerance 1e-4), equality of the gradient across ranks, and the time of that 5-shot gradient."""
    import torch
    import torch.distributed as dist
    from red_diffeq_b200 import ShardedFWIForward, s_normalize_none, v_denormalize
    from red_diffeq_b200.utils import synthetic
    path = os.path.join(ROOT, "tests", "golden", "marmousi.npz")
    if not os.path.exists(path):
        return {"unavailable": "tests/golden/marmousi.npz missing"}
    z = np.load(path, allow_pickle=False)
    ctx = json.loads(str(z["ctx"]))
    for k in ("n_grid", "nt", "nbc", "ng", "ns"):
        ctx[k] = int(ctx[k])
    dev = env.dev
    op = ShardedFWIForward(dict(ctx), dev, mode="shots", normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    v = torch.tensor(z["v"], device=dev, requires_grad=True)
    cot = torch.tensor(synthetic.cotangent((1, ctx["ns"], ctx["nt"], ctx["ng"]), seed=int(z["cot_seed"])), device=dev)
    _, _, shots = op.partition(1)
    cot_local = cot[:, torch.as_tensor(shots, device=dev, dtype=torch.long)] if len(shots) else None

    def grad_once():
        v.grad = None
        seis = op(v)
        if cot_local is None:
            seis.sum().backward()      # idle rank: an empty result that still reaches the all-reduce
        else:
            seis.backward(cot_local)
        return v.grad

    g = grad_once().clone()
    ref = torch.tensor(z["grad_f32"], device=dev)
    rel = float(((g - ref).double().norm() / ref.double().norm()).item())
    # bit-equality across ranks: every rank compares with rank 0's copy
    g0 = g.clone()
    if env.world > 1:
        dist.broadcast(g0, src=0)
    differs = float((g != g0).any().item())
    differs = env.max_over_ranks([differs])[0]
    # time of the reference's real shot count, sharded (each rank runs the few-shot cluster configuration)
    for _ in range(3):
        grad_once()
    env.barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    reps = 10
    for _ in range(reps):
        grad_once()
    a1.record()
    env.barrier()
    ms = env.max_over_ranks([a0.elapsed_time(a1) / reps])[0]
    from red_diffeq_b200.solvers.sharding import split_range
    counts = [hi - lo for lo, hi in (split_range(ctx["ns"], env.world, r) for r in range(env.world))]
    for o in op._ops.values():
        o.release_memory()
    return {"grad_rel_l2_vs_fixture": rel, "tolerance": 1e-4, "ranks_bit_identical": differs == 0.0,
            "fixture": "tests/golden/marmousi.npz (reference pde.py autograd, ns = 5, nt = 1000)", "shots_per_rank": counts,
            "ns5_ms_per_gradient": ms,
            "ns5_pairs_per_s": ctx["ns"] * (70 + 2 * ctx["nbc"]) * (190 + 2 * ctx["nbc"]) * ctx["nt"] / (ms * 1e-3)}


def sharded_block(env, args):
    """Strong scaling of the shot-sharded Marmousi gradient (north-star: ">= 85 % scaling efficiency to 8 GPUs on shot-sharded
    Marmousi inversion"): 176 shots, ShardedFWIForward, one NCCL all-reduce of d loss / d v inside every timed step."""
    steps, warmup = max(3, min(args.steps, 10)), max(3, min(args.warmup, 3))
    m = measure_workload(env, "marmousi_sharded", steps, warmup, opts=args.opt)
    parity = sharded_parity(env)
    pairs = m["pairs_total"]
    p = m["plan"]
    return {"workload": "marmousi_sharded: 1 Marmousi-shaped model (70x190, padded 310x430) x 176 shots x 1000 levels, shots dealt over the ranks",
            "scaling": "strong", "n_gpus": env.world, "steps": steps, "warmup": warmup,
            "ms_per_step": m["ms_per_step"], "value": pairs / (m["ms_per_step"] * 1e-3), "unit": "pairs/s",
            "e2e": {"value": pairs / (m["e2e_ms_per_step"] * 1e-3), "unit": "pairs/s", "ms_per_step": m["e2e_ms_per_step"],
                    "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"]},
            "allreduce_us": m["allreduce_us"], "allreduce_bytes": 70 * 190 * 4,
            "shots_on_rank0": m["ns_local"], "phase_ms": {"forward": m["fwd_ms"], "adjoint_incl_allreduce": m["adj_ms"]},
            "kernels_ms": {k: v * 1e-3 for k, v in m["kernel_us"].items() if v > 0},
            "engine": config_block(m)["engine"], "cluster_wave": p["cluster_wave"],
            "limits": "per rank 176/N shots run as ceil(shots / %d) rounds of co-resident clusters (wave quantisation) at a fixed "
                      "per-level latency; the all-reduce is %d B" % (p["cluster_wave"], 70 * 190 * 4),
            "parity": parity}


# ---------------------------------------------------------------------------------------------------------------------
def red_iter_block(env, args):
    """Seconds per FULL RED-DiffEq iteration (BASELINE.json metric, second half; configs[1] as written): FD forward + adjoint
    of the B200 operator, the reference's own U-Net + GaussianDiffusion as the RED regulariser (unmodified
    red_diffeq/models/diffusion.py staged under baseline/_ref, sizes of configs/*/red-diffeq.yaml, RANDOM-INIT -- the weights
    are not in the repository, and the reference itself falls back to random init, scripts/run_inversion.py:68-70), Adam,
    clamp, cosine schedule and the per-iteration metrics, through InversionEngine.optimize.  Two regulariser call patterns:
    the reference's RED_DiffEq (autograd graph recorded and discarded, one U-Net call per patch) and this repo's REDDiffEq
    (no_grad, patches batched)."""
    import torch
    try:
        from baseline import ref_loader
        if not ref_loader.available():
            return {"unavailable": "baseline/_ref is empty (run baseline/stage_reference.py where /root/reference exists)"}
        from red_diffeq_b200 import FWIForward, InversionEngine, REDDiffEq, s_normalize_none, v_denormalize
        from red_diffeq_b200.utils import synthetic
        dev = env.dev
        dm = ref_loader.build_diffusion(dev)
        base, inv_mod, ssim_mod = ref_loader.load("red_diffeq.regularization.base", "red_diffeq.core.inversion", "red_diffeq.utils.ssim")
        out = {"denoiser": "reference Unet(dim=64, dim_mults=(1,2,4,8), channels=1) + GaussianDiffusion(image_size=72, timesteps=1000, "
                           "objective=pred_noise), %.1f M parameters, random init, fp32, eval" % (sum(p.numel() for p in dm.parameters()) / 1e6),
               "loop": "red_diffeq_b200.InversionEngine.optimize (fused misfit, metrics fetched once), lr 0.03, reg_lambda 0.75, sigma_x0 1e-4"}
        for tag, kind, nz, nx, B, ts in (("openfwi_b64", "openfwi", 70, 70, 64, 6), ("marmousi_b1", "marmousi", 70, 190, 1, 20)):
            ctx = make_ctx(kind)
            op = FWIForward(dict(ctx), dev, normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
            for kv in args.opt:
                k, v = kv.split("=")
                op.set_option(k, int(v))
            mu_true_n = torch.tensor(synthetic.velocity_models(B, nz, nx, seed=synthetic.SEED), device=dev)
            with torch.no_grad():
                y = op(mu_true_n)
            mu0 = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(mu_true_n, (5, 5, 5, 5), mode="replicate"), 11, stride=1)
            mu0 = torch.nn.functional.pad(mu0, (1, 1, 1, 1), value=0.0)                  # scripts/run_inversion.py:156
            mu_true = v_denormalize(mu_true_n)
            pairs = op.pairs_per_gradient(B, nz, nx)
            ref_method = base.RegularizationMethod("diffusion", dm)
            ours = REDDiffEq(dm)

            def time_reg(fn):
                mu = mu0.clone().requires_grad_(True)
                for _ in range(2):
                    fn(mu)[0].sum().backward()
                torch.cuda.synchronize(dev)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(5):
                    mu.grad = None
                    fn(mu)[0].sum().backward()
                a1.record()
                torch.cuda.synchronize(dev)
                return a0.elapsed_time(a1) / 5

            def time_solver():
                mask = torch.ones_like(y)
                for _ in range(2):
                    v = mu0[:, :, 1:-1, 1:-1].clone().requires_grad_(True)
                    op.misfit(v, y, mask).sum().backward()
                torch.cuda.synchronize(dev)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(5):
                    v = mu0[:, :, 1:-1, 1:-1].clone().requires_grad_(True)
                    op.misfit(v, y, mask).sum().backward()
                a1.record()
                torch.cuda.synchronize(dev)
                return a0.elapsed_time(a1) / 5

            def time_loop(engine):
                engine.optimize(mu0, mu_true, y, op, ts=2, lr=0.03, reg_lambda=0.75, regularization="diffusion")   # warm-up
                _, res = engine.optimize(mu0, mu_true, y, op, ts=ts, lr=0.03, reg_lambda=0.75, regularization="diffusion")
                return engine.last_loop_seconds / ts, res

            reg_ms_ref = time_reg(lambda mu: ref_method.get_reg_loss(mu))
            reg_ms_ours = time_reg(lambda mu: ours(mu))
            solver_ms = time_solver()
            s_ref, _ = time_loop(InversionEngine(regularization="diffusion", regularizer=lambda mu: ref_method.get_reg_loss(mu), cuda_graph=False))
            s_eager, res = time_loop(InversionEngine(dm, regularization="diffusion", cuda_graph=False, overlap_regularizer=False))
            blk = {"models": B, "shots": ctx["ns"], "nt": ctx["nt"], "iterations_timed": ts,
                   "s_per_iter_reference_pattern": s_ref,           # reference's RED_DiffEq calls inside this repo's loop, eager
                   "s_per_iter_eager": s_eager,                     # REDDiffEq (no_grad, batched patches), eager, one stream
                   "solver_ms": solver_ms, "reg_ms_reference_pattern": reg_ms_ref, "reg_ms_ours": reg_ms_ours}
            try:   # the engine's default for the diffusion regulariser: one CUDA graph per iteration, U-Net on a second stream
                eng = InversionEngine(dm, regularization="diffusion")
                s_default, res = time_loop(eng)
                blk["s_per_iter"] = s_default
                blk["cuda_graph"] = bool(eng.used_cuda_graph)
            except Exception as e:
                blk["s_per_iter"] = s_eager
                blk["cuda_graph"] = False
                blk["cuda_graph_error"] = f"{type(e).__name__}: {str(e)[:200]}"
                torch.cuda.synchronize(dev)
            try:   # the reference's OWN loop (unmodified core/inversion.py: its losses, SSIM per model, six host round trips per
                   # iteration) with this repo's operator dropped in for fwi_forward -- the north-star's drop-in, end to end
                os.environ.setdefault("TQDM_DISABLE", "1")   # its progress bar would flood stderr
                their = inv_mod.InversionEngine(dm, ssim_mod.SSIM(window_size=11), regularization="diffusion", sigma_x0=1e-4)
                their.optimize(mu0, mu_true, y, op, ts=2, lr=0.03, reg_lambda=0.75, regularization="diffusion")
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                their.optimize(mu0, mu_true, y, op, ts=ts, lr=0.03, reg_lambda=0.75, regularization="diffusion")
                torch.cuda.synchronize(dev)
                blk["s_per_iter_reference_loop_with_our_operator"] = (time.perf_counter() - t0) / ts
            except Exception as e:
                blk["s_per_iter_reference_loop_with_our_operator"] = None
                blk["reference_loop_error"] = f"{type(e).__name__}: {str(e)[:200]}"
            blk["solver_share"] = solver_ms * 1e-3 / blk["s_per_iter"]
            blk["pairs_per_s_through_the_iteration"] = pairs / blk["s_per_iter"]
            blk["obs_loss_first_last"] = [float(res[0]["obs_losses"][0]), float(res[0]["obs_losses"][-1])]
            out[tag] = blk
            op.release_memory()
            del op, y
            torch.cuda.empty_cache()
        # headline keys of the block = configs[1] (64 OpenFWI models)
        h = out["openfwi_b64"]
        out.update({"s_per_iter": h["s_per_iter"], "solver_share": h["solver_share"],
                    "reg_ms_reference_pattern": h["reg_ms_reference_pattern"], "reg_ms_ours": h["reg_ms_ours"],
                    "published_reference": "2.24-2.25 s / iteration for ONE OpenFWI model on an RTX 3090 (example/example_openfwi.ipynb:655-657, BASELINE.md 1)"})
        return out
    except Exception as e:
        import traceback
        traceback.print_exc(file=sys.stderr)
        try:
            torch.cuda.synchronize(env.dev)
            torch.cuda.empty_cache()
        except Exception:
            pass
        return {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="openfwi_b64", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override models per GPU (debugging)")
    ap.add_argument("--nt", type=int, default=0, help="override time levels (debugging; invalidates the headline)")
    ap.add_argument("--ns", type=int, default=0, help="override shots per model (debugging; invalidates the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-pde", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the shot-sharded Marmousi block")
    ap.add_argument("--no-red-iter", action="store_true", help="skip the full RED-DiffEq iteration block")
    ap.add_argument("--only-headline", action="store_true", help="the headline workload only (profiling runs)")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (e.g. chunk_models=8)")
    ap.add_argument("--history-segment", type=int, default=None,
                    help="wavefield history policy: 0 = every level, K >= 3 = checkpoint every K levels (default: automatic)")
    args = ap.parse_args()
    if args.only_headline:
        args.no_cpu_baseline = args.no_reference_pde = args.no_sharded = args.no_red_iter = True

    if args.impl == "reference":
        run_reference_arm(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    env = Env()
    if env.world > 1:
        dist.init_process_group("nccl", device_id=env.dev)
    rank, world = env.rank, env.world

    sampler = ClockSampler(env.local_rank) if rank == 0 else None
    m = measure_workload(env, args.workload, args.steps, args.warmup, opts=args.opt, history_segment=args.history_segment,
                         batch=args.batch, nt=args.nt, ns=args.ns, sampler=sampler)
    sharded = args.workload == "marmousi_sharded"
    line = None
    if rank == 0:
        value = m["pairs_total"] / (m["ms_per_step"] * 1e-3)
        line = {
            "metric": "FD cell-updates/s (fwd+adjoint)", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(m),
            "e2e": {"value": m["pairs_total"] / (m["e2e_ms_per_step"] * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"]},
            "gpu_launches": int((m["launches_f"] + m["launches_b"]) * args.steps),
            "roofline": roofline_block(m, args.steps),
            "phase_ms": {"forward": m["fwd_ms"], "adjoint": m["adj_ms"]},
            "clocks": m["clocks"],
        }
    # ---- the north-star's multi-GPU path, at every N (N = 1 is the denominator of the scaling efficiency) ----------------
    if not args.no_sharded and not sharded:
        try:
            blk = sharded_block(env, args)
        except Exception as e:
            import traceback
            traceback.print_exc(file=sys.stderr)
            blk = {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}
        if rank == 0:
            line["sharded"] = blk
    if world > 1:
        env.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if world == 1 and not args.no_red_iter:
            line["red_iter"] = red_iter_block(env, args)
        kind, nz, nx, _ = WORKLOADS[args.workload]
        if not args.no_cpu_baseline:     # rank 0 alone, after the collectives are done (any N)
            base, _, _ = cpu_baseline(make_ctx(kind), nz, nx)
            line["cpu_baseline"] = base
        if not args.no_reference_pde and world == 1:
            line["reference_pde"] = reference_pde_leg()
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
