"""BASELINE config 5: stencil scaling sweep -- interior n x n (nbc = 120), ns shots, forward+adjoint pairs/s.

    python tools/sweep.py [--out profiles/sweep_r2.md] [--quick] [--min-n N] [--max-n N] [--min-shots S]
    torchrun --nproc-per-node N --master-addr 127.0.0.1 ... tools/sweep.py ...      (N GPUs)

One GPU: the operator alone.  N GPUs (one rank per GPU, NCCL): the SAME case with its shots dealt over the ranks by
ShardedFWIForward -- strong scaling, the gradient all-reduce inside the timed region, time = max over ranks; cases with
fewer shots than ranks are skipped.  Per case: the engine the library picks, the history policy, ms for forward and adjoint,
pairs/s (whole job) and the fraction of N x the 28 B/pair HBM roofline (MEASURED_PEAKS.json).  nt = 1000 unless stated (the
largest grids use fewer levels so that a case takes seconds; stated in the table).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from red_diffeq_b200 import FWIForward, ShardedFWIForward, s_normalize_none, v_denormalize  # noqa: E402
from red_diffeq_b200.utils import synthetic  # noqa: E402

OPTS = {}
CASES = [  # (n, ns, B, nt)
    (70, 5, 1, 1000), (70, 5, 64, 1000), (128, 1, 1, 1000), (128, 16, 1, 1000), (128, 64, 1, 1000), (128, 256, 1, 1000),
    (256, 16, 1, 1000), (256, 64, 1, 1000), (256, 256, 1, 1000), (512, 4, 1, 1000), (512, 16, 1, 1000), (512, 64, 1, 1000),
    (1024, 4, 1, 1000), (1024, 16, 1, 1000), (2048, 4, 1, 500), (2048, 16, 1, 500), (4096, 1, 1, 300), (4096, 4, 1, 300),
    (4096, 16, 1, 150),
]


def run_case(n, ns, B, nt, peak, dev, rank, world):
    ctx = dict(n_grid=n, nt=nt, dx=10.0, dt=0.001, nbc=120, f=15.0, sz=10, gz=10, ng=n, ns=ns)
    kw = dict(normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    if world > 1:
        wrapper = ShardedFWIForward(dict(ctx), dev, mode="shots" if B < world else "auto", **kw)
        mode, models, shots = wrapper.partition(B)
        op = wrapper._operator(shots)
        fwd, nb, ns_local = wrapper, models.stop - models.start, len(shots)
    else:
        op = FWIForward(ctx, dev, **kw)
        fwd, nb, ns_local = op, B, ns
    for k, val in OPTS.items():
        op.set_option(k, val)
    v = torch.tensor(synthetic.velocity_models(B, n, n), device=dev)
    plan = op._plan_for(n, n, dev)
    cot = torch.ones((nb, ns_local, nt, n), device=dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        vv = v.detach().requires_grad_(True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); s = fwd(vv); e[1].record(); s.backward(cot); e[2].record()
        return e
    step(); step(); sync()
    f_ms = a_ms = None
    for _ in range(2):  # best of two timed evaluations (each the max over ranks)
        e = step(); sync()
        t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        f, a = t.tolist()
        if f_ms is None or f + a < f_ms + a_ms:
            f_ms, a_ms = f, a
    pairs = B * ns * (n + 240) ** 2 * nt
    rate = pairs / ((f_ms + a_ms) * 1e-3)
    seg = plan.get("history_segment")
    clustered = plan.get("cluster_size_used") and OPTS.get("engine") != 1
    sp = plan.get("adj_split")
    eng_f = "cluster C=%d R=%d" % (plan.get("cluster_size_last"), plan.get("cluster_rows_last")) if clustered and (not seg or sp in (2, 5)) else "per-level"
    eng_a = {1: "cluster split", 2: "cluster split, forward recomputed", 4: "cluster resident (TMEM imaging)", 5: "cluster resident, forward recomputed", 3: "per-level split (chunks of %d shots)" % plan.get("u_chunk_used")}.get(sp) or "per-level fused"
    op.release_memory()
    hist = "none (recomputed)" if sp in (2, 5) else ("checkpoint K=%d" % seg if seg else "full")
    return dict(n=n, ns=ns, B=B, nt=nt, forward_ms=f_ms, adjoint_ms=a_ms, pairs_per_s=rate, frac=rate * 28 / (world * peak * 1e9),
                engine_fwd=eng_f, engine_adj=eng_a, history=hist, gpus=world, shots_per_rank=ns_local)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--min-n", type=int, default=0, help="only cases with interior n >= this")
    ap.add_argument("--max-n", type=int, default=1 << 30, help="only cases with interior n <= this")
    ap.add_argument("--min-shots", type=int, default=0, help="only cases with at least this many shots")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rows = []
    global OPTS
    OPTS = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in args.opt)
    cases = CASES[:4] if args.quick else [c for c in CASES if args.min_n <= c[0] <= args.max_n and c[1] >= args.min_shots]
    for case in cases:
        if world > 1 and case[1] * case[2] < world:
            continue  # fewer work items than ranks
        try:
            rows.append(run_case(*case, peak, dev, rank, world))
        except Exception as ex:  # report and continue
            rows.append(dict(n=case[0], ns=case[1], B=case[2], nt=case[3], error=str(ex)[:120]))
            if world > 1:
                raise
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
        torch.cuda.empty_cache()
    if rank == 0:
        lines = ["| interior n | shots | models | nt | GPUs | engine fwd / adj | history | fwd ms | adj ms | pairs/s (whole job) | of GPUs x 28 B/pair HBM roofline (%.0f GB/s each) |" % peak,
                 "|---|---|---|---|---|---|---|---|---|---|---|"]
        for r in rows:
            if "error" in r:
                lines.append(f"| {r['n']} | {r['ns']} | {r['B']} | {r['nt']} | {world} | error: {r['error']} |||||| ")
            else:
                lines.append(f"| {r['n']} | {r['ns']} | {r['B']} | {r['nt']} | {r['gpus']} | {r['engine_fwd']} / {r['engine_adj']} | {r['history']} | "
                             f"{r['forward_ms']:.1f} | {r['adjoint_ms']:.1f} | {r['pairs_per_s']:.3e} | {r['frac']:.2f} |")
        text = "\n".join(lines)
        print(text)
        if args.out:
            with open(args.out, "w") as f:
                f.write("# Stencil scaling sweep (BASELINE config 5), %d B200, forward + adjoint\n\n" % world + text + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
