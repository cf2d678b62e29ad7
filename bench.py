#!/usr/bin/env python
"""bench.py -- FD forward+adjoint throughput of the B200 path (and of the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one gradient evaluation of the hot path over one batch of synthetic velocity models:
forward modelling of every shot (nt levels) + the reverse-time adjoint with the imaging condition,
i.e. B*ns*nzp*nxp*nt forward+adjoint cell-update *pairs* (SURVEY.md 8d).  Default workload = BASELINE.json
configs[1] (OpenFWI 70x70, 64 models x 5 shots x 1000 levels) on one GPU; with N GPUs every rank runs its
own 64 models (models are independent: weak scaling, no data-path collective).

Prints ONE JSON line (see the keys at the bottom).  `value` is measured with inputs resident in HBM,
`e2e` through the public operator (FWIForward + autograd) from pinned host buffers.
"""
import argparse
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
}

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the `ncu --set full`
# capture of the same workload committed under profiles/ (None = not captured for this workload)
NCU_TRAFFIC_BYTES = {
    # per launch, from profiles/launches_r1_v2_b64.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on
    # the bench command): the forward launch writes the 123.7 GB history; each of the 5 adjoint-field launches writes a
    # 64-shot u history; each imaging launch reads both histories of its 64 shots
    ("openfwi_b64", "forward"): 123.8e9,
    ("openfwi_b64", "adjoint_field"): 24.7e9,
    ("openfwi_b64", "imaging"): 52.0e9,      # profiles/launches_r1_v2_b64.csv: 51.9 GB read + 0.05 GB written per launch
    # fused cluster adjoint (adj_mode=1), profiles/ncu_adj_cluster_r1_full_b64.txt
    ("openfwi_b64", "adjoint_loop"): 124.13e9,
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
    return ctx


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax = float(r[1])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # only samples under load (top half) describe the timed region
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(sm_load)) if sm_load else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(ctx, nz, nx, target_seconds=20.0):
    """Times the oracle port (CPU restatement of the reference algorithm) on a bounded sample of the workload."""
    from oracle import build_oracle, fwi_oracle
    from red_diffeq_b200.utils import synthetic
    build_oracle.build()
    # all the host cores this process may use: launchers such as torchrun export OMP_NUM_THREADS=1 to their workers, which
    # would time the CPU arm on one thread at N > 1
    try:
        fwi_oracle.set_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        fwi_oracle.set_threads(os.cpu_count() or 1)
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


def run_reference_arm(args, rank, world):
    """--impl reference: the reference algorithm on the host CPU (oracle port; the Python reference cannot travel)."""
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
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (e.g. chunk_models=8)")
    ap.add_argument("--history-segment", type=int, default=None,
                    help="wavefield history policy: 0 = every level, K >= 3 = checkpoint every K levels (default: automatic)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from red_diffeq_b200 import FWIForward, ShardedFWIForward, s_normalize_none, v_denormalize
    from red_diffeq_b200.utils import synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kind, nz, nx, B = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    ctx = make_ctx(kind)
    if args.nt:
        ctx["nt"] = args.nt
    if args.ns:
        ctx["ns"] = args.ns
    sharded = args.workload == "marmousi_sharded"
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
    for kv in args.opt:
        k, v = kv.split("=")
        op.set_option(k, int(v))
    if args.history_segment is not None:
        op.set_history_segment(args.history_segment)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fwd = wrapper if sharded else op   # the sharded wrapper adds the gradient all-reduce to backward()

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
    for _ in range(args.warmup):
        step_resident()
    barrier()
    op._plan_for(nz, nx, dev).set("timing", 1)   # library-side CUDA events per kernel class, inside the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_stop = torch.cuda.Event(enable_timing=True)
    t_start.record()
    evs = [step_resident() for _ in range(args.steps)]
    t_stop.record()
    barrier()
    elapsed_ms = t_start.elapsed_time(t_stop)
    kernel_us = {k: op._plan_for(nz, nx, dev).get("us_" + k) for k in ("forward", "adjoint_field", "imaging", "adjoint_loop")}
    kernel_n = {k: op._plan_for(nz, nx, dev).get("n_" + k) for k in ("forward", "adjoint_field", "imaging", "adjoint_loop")}
    op._plan_for(nz, nx, dev).set("timing", 0)
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    adj_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    launches_f, launches_b = evs[-1][3], evs[-1][4]

    # ---- end-to-end timing from pinned host buffers (e2e) -------------------------------------------
    for _ in range(min(args.warmup, 1)):
        step_e2e()
    barrier()
    e_start = torch.cuda.Event(enable_timing=True)
    e_stop = torch.cuda.Event(enable_timing=True)
    e_start.record()
    for _ in range(args.steps):
        step_e2e()
    e_stop.record()
    barrier()
    e2e_ms = e_start.elapsed_time(e_stop)
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = t.tolist()
        tot = torch.tensor([float(pairs_rank)], device=dev, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        pairs_total = tot.item()
    else:
        pairs_total = float(pairs_rank)

    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        value = pairs_total / (ms_per_step * 1e-3)
        e2e_value = pairs_total / (e2e_ms / args.steps * 1e-3)
        peak, peak_src = measured_peak_gbs()
        # Per-kernel-class device times come from CUDA events the library records on the launch stream around
        # each class of launches inside the timed region (rdfwi_plan_set "timing").
        plan = op._plan_for(nz, nx, dev)
        eng = op.options.get("engine", 0)
        seg = plan.get("history_segment")
        recompute = plan.get("adj_split") == 2      # no history kept: the backward pass re-runs the forward kernel per chunk
        fwd_cluster = eng != 1 and plan.get("cluster_size_used") > 0 and (seg == 0 or recompute)
        adj_split = plan.get("adj_split") >= 1
        adj_cluster = (not adj_split) and eng != 1 and plan.get("adj_cluster_size_used") > 0 and seg == 0
        us = {k: kernel_us[k] / args.steps for k in kernel_us}
        n = {k: kernel_n[k] // args.steps for k in kernel_n}
        cell_updates = float(cells_level) * nt
        kernels = {
            "forward": {"kernel": "k_fwd_cluster<EXACT>" if fwd_cluster else "k_step_tile<EXACT>", "us": us["forward"],
                        "launches": n["forward"] if fwd_cluster else launches_f - 3,
                        # recompute tier: the forward kernel really runs twice per step (modelling + per-chunk recompute)
                        "algo_bytes": ALGO_BYTES_FWD * cell_updates * (2 if recompute else 1)},
        }
        if adj_split:
            # the adjoint's 16 B / cell-update split as: adjoint-field kernel (read u_{t+1}, u_{t+2}, write u_t = 12 B; all of
            # it stays in shared memory, only the 4 B history write reaches HBM) + imaging kernel (pointwise: read p_t and
            # u_t once each = 8 B, which is exactly what it streams from HBM)
            tiled = plan.get("adj_split") == 3   # per-level engine: nt tiled launches per chunk of shots
            kernels["adjoint_field"] = {"kernel": "k_step_tile<ADJ>" if tiled else "k_fwd_cluster<ADJ>", "us": us["adjoint_field"],
                                        "launches": n["adjoint_field"] * (nt if tiled else 1), "algo_bytes": 12.0 * cell_updates}
            kernels["imaging"] = {"kernel": "k_imaging", "us": us["imaging"], "launches": n["imaging"],
                                  "algo_bytes": 8.0 * cell_updates}
        else:
            kernels["adjoint_loop"] = {"kernel": "k_adj_cluster" if adj_cluster else "k_adj_step", "us": us["adjoint_loop"],
                                       "launches": n["adjoint_loop"] if adj_cluster else launches_b - 7,
                                       "algo_bytes": ALGO_BYTES_ADJ * cell_updates}
        peak, peak_src = measured_peak_gbs()
        for kk in kernels.values():
            kk["achieved_gbs"] = kk["algo_bytes"] / (kk["us"] * 1e-6) / 1e9 if kk["us"] > 0 else None
            kk["avg_launch_us"] = kk["us"] / max(kk["launches"], 1)
        for name, kk in kernels.items():
            kk["frac"] = kk["achieved_gbs"] / peak if kk["achieved_gbs"] else None
            # what really crossed the HBM pins (ncu dram bytes of one launch x launches), where a capture exists: the
            # cluster-resident kernels move far less than their algorithmic bytes, the imaging kernel exactly its own
            tb = NCU_TRAFFIC_BYTES.get((args.workload, name))
            kk["dram_gbs"] = tb * kk["launches"] / (kk["us"] * 1e-6) / 1e9 if tb and kk["us"] > 0 else None
        dom_key = max(kernels, key=lambda q: kernels[q]["us"])
        dom = kernels[dom_key]
        fwd_achieved = ALGO_BYTES_FWD * cell_updates / (fwd_ms * 1e-3) / 1e9
        adj_achieved = ALGO_BYTES_ADJ * cell_updates / (adj_ms * 1e-3) / 1e9
        line = {
            "metric": "FD cell-updates/s (fwd+adjoint)", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "models_per_gpu": B, "shots_per_model": ns,
                       "shots_on_rank0": ns_local, "nt": nt,
                       "padded_grid": [nzp, nxp], "pairs_per_step_per_gpu": pairs_rank,
                       "l2_policy": "working set (wavefield histories, %.1f GB streamed per step) far exceeds the 126 MB L2; no flush needed"
                                    % ((plan.history_bytes(B, seg) + plan.workspace_bytes(B)) / 1e9),
                       "history": ("none kept: forward field recomputed per chunk of %d shots in the backward pass" % plan.get("u_chunk_used")) if recompute
                                  else (("checkpoint pairs every %d levels" % seg) if seg else "every level"),
                       "engine": {"forward": "cluster-resident (C=%d, %d rows per thread)" % (plan.get("cluster_size_last"), plan.get("cluster_rows_last")) if fwd_cluster else "per-level",
                                  "adjoint": ("split: per-level tiled adjoint field + streaming imaging" if plan.get("adj_split") == 3 else
                                              "split: cluster-resident adjoint field (C=%d) + streaming imaging" % plan.get("cluster_size_last")) if adj_split
                                  else ("cluster-resident fused (C=%d)" % plan.get("adj_cluster_size_used") if adj_cluster else "per-level")},
                       "options": dict(op.options)},
            "e2e": {"value": e2e_value, "unit": "pairs/s",
                    "h2d_bytes_per_step": int(vn_host.numel() * 4 + y_host.numel() * 4),
                    "d2h_bytes_per_step": int(grad_host.numel() * 4 + loss_host.numel() * 4)},
            "gpu_launches": int((launches_f + launches_b) * args.steps),
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                         "frac": dom["frac"], "frac_of_nominal_8TBs": dom["achieved_gbs"] / 8000.0 if dom["achieved_gbs"] else None,
                         "traffic": NCU_TRAFFIC_BYTES.get((args.workload, dom_key)),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": dom["algo_bytes"] / max(dom["launches"], 1),
                         "avg_launch_us": dom["avg_launch_us"], "launches_per_step": dom["launches"],
                         "share_of_step": dom["us"] * 1e-3 / ms_per_step,
                         "kernels": {k: {"kernel": v["kernel"], "ms_per_step": v["us"] * 1e-3, "launches_per_step": v["launches"],
                                         "algorithmic_GB_per_step": v["algo_bytes"] / 1e9, "achieved_gbs": v["achieved_gbs"],
                                         "frac": v["frac"], "traffic": NCU_TRAFFIC_BYTES.get((args.workload, k)),
                                         "measured_dram_gbs": v["dram_gbs"],
                                         "measured_dram_frac": v["dram_gbs"] / peak if v["dram_gbs"] else None}
                                     for k, v in kernels.items()},
                         # SURVEY.md 8(d) aggregates: 12 B forward, 16 B adjoint (both adjoint kernels together), 28 B pair
                         "forward_frac": fwd_achieved / peak, "adjoint_frac": adj_achieved / peak,
                         "pair_frac": (ALGO_BYTES_PAIR * pairs_rank / ((fwd_ms + adj_ms) * 1e-3) / 1e9) / peak},
            "phase_ms": {"forward": fwd_ms, "adjoint": adj_ms},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            base, _, _ = cpu_baseline(ctx, nz, nx)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
