"""BASELINE config 5: stencil scaling sweep -- interior n x n (nbc = 120), ns shots, forward+adjoint pairs/s on one GPU.

    python tools/sweep.py [--out profiles/sweep_r1.md] [--quick]

Per case: the engine the library picks, whether the history was checkpointed, ms for forward and adjoint, pairs/s and the
fraction of the 28 B/pair HBM roofline (MEASURED_PEAKS.json).  nt = 1000 unless stated (largest grids use fewer levels so
that a case takes seconds).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from red_diffeq_b200 import FWIForward, s_normalize_none, v_denormalize  # noqa: E402
from red_diffeq_b200.utils import synthetic  # noqa: E402

OPTS = {}
CASES = [  # (n, ns, B, nt)
    (70, 5, 1, 1000), (70, 5, 64, 1000), (128, 1, 1, 1000), (128, 16, 1, 1000), (128, 64, 1, 1000), (128, 256, 1, 1000),
    (256, 16, 1, 1000), (256, 64, 1, 1000), (512, 4, 1, 1000), (512, 16, 1, 1000), (512, 64, 1, 1000),
    (1024, 4, 1, 1000), (1024, 16, 1, 1000), (2048, 4, 1, 500), (2048, 16, 1, 500), (4096, 1, 1, 300), (4096, 4, 1, 300),
]


def run_case(n, ns, B, nt, peak):
    ctx = dict(n_grid=n, nt=nt, dx=10.0, dt=0.001, nbc=120, f=15.0, sz=10, gz=10, ng=n, ns=ns)
    op = FWIForward(ctx, "cuda:0", normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    for k, val in OPTS.items():
        op.set_option(k, val)
    v = torch.tensor(synthetic.velocity_models(B, n, n), device="cuda:0")
    plan = op._plan_for(n, n, torch.device("cuda:0"))

    def step():
        vv = v.detach().requires_grad_(True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); s = op(vv); e[1].record(); s.backward(torch.ones_like(s)); e[2].record()
        return e
    step(); step(); torch.cuda.synchronize()
    f_ms = a_ms = None
    for _ in range(2):  # best of two timed evaluations
        e = step(); torch.cuda.synchronize()
        f, a = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        if f_ms is None or f + a < f_ms + a_ms:
            f_ms, a_ms = f, a
    pairs = B * ns * (n + 240) ** 2 * nt
    rate = pairs / ((f_ms + a_ms) * 1e-3)
    seg = plan.get("history_segment")
    eng_f = "cluster C=%d R=%d" % (plan.get("cluster_size_last"), plan.get("cluster_rows_last")) if plan.get("cluster_size_used") and not seg and OPTS.get("engine") != 1 else "per-level"
    sp = plan.get("adj_split")
    eng_a = {1: "cluster split", 2: "cluster split, forward recomputed", 3: "per-level split"}.get(sp) or \
        (("cluster fused C=%d" % plan.get("adj_cluster_size_used")) if plan.get("adj_cluster_size_used") and not seg and OPTS.get("engine") != 1 else "per-level fused")
    op.release_memory()
    return dict(n=n, ns=ns, B=B, nt=nt, forward_ms=f_ms, adjoint_ms=a_ms, pairs_per_s=rate, frac=rate * 28 / (peak * 1e9),
                engine_fwd=eng_f, engine_adj=eng_a, history="checkpoint K=%d" % seg if seg else "full")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--min-n", type=int, default=0, help="only cases with interior n >= this")
    ap.add_argument("--max-n", type=int, default=1 << 30, help="only cases with interior n <= this")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value")
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rows = []
    global OPTS
    OPTS = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in args.opt)
    for case in (CASES[:4] if args.quick else [c for c in CASES if args.min_n <= c[0] <= args.max_n]):
        try:
            rows.append(run_case(*case, peak))
        except Exception as ex:  # report and continue
            rows.append(dict(n=case[0], ns=case[1], B=case[2], nt=case[3], error=str(ex)[:120]))
        print(rows[-1], flush=True)
        torch.cuda.empty_cache()
    lines = ["| interior n | shots | models | nt | engine fwd / adj | history | fwd ms | adj ms | pairs/s | of 28 B/pair HBM roofline (%.0f GB/s) |" % peak,
             "|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        if "error" in r:
            lines.append(f"| {r['n']} | {r['ns']} | {r['B']} | {r['nt']} | error: {r['error']} |||||| ")
        else:
            lines.append(f"| {r['n']} | {r['ns']} | {r['B']} | {r['nt']} | {r['engine_fwd']} / {r['engine_adj']} | {r['history']} | "
                         f"{r['forward_ms']:.1f} | {r['adjoint_ms']:.1f} | {r['pairs_per_s']:.3e} | {r['frac']:.2f} |")
    text = "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write("# Stencil scaling sweep (BASELINE config 5), one B200, forward + adjoint\n\n" + text + "\n")


if __name__ == "__main__":
    main()
