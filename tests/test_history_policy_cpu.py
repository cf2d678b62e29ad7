"""CPU: the wavefield-history policy of the operator (FWIForward._choose_segment) exercised with a fake plan and
a fake amount of free HBM -- which tier is chosen when, that a rejected tier's options are undone, and that the decision is
cached per (plan, batch).  No GPU call is made: torch.cuda's memory queries are monkeypatched."""
import math

import pytest

torch = pytest.importorskip("torch")

GB = 1e9


class FakePlan:
    """Answers the size queries the policy asks, from the same rules as the library (rdfwi_history_bytes, carve())."""

    def __init__(self, nt=1000, ns=5, level_floats=96720, clustered=True, wave=33):
        self.nt, self.ns, self._level, self.clustered, self.wave = nt, ns, level_floats, clustered, wave
        self.opts = {"history_segment": 0, "adj_mode": 0, "u_chunk_shots": 0, "scratch_mb": 0, "imaging": 0}
        self.log = []
        self.size_queries = 0

    def level_floats(self):
        return self._level

    def set(self, key, value):
        self.opts[key] = int(value)
        self.log.append((key, int(value)))

    def get(self, key):
        if key == "cluster_size_used":
            return 4 if self.clustered else 0
        if key == "cluster_wave":
            return self.wave
        return self.opts[key]

    def history_bytes(self, B, segment):
        per_level = 4.0 * self._level * B * self.ns
        if segment == 0:
            return per_level * self.nt
        if segment >= self.nt:
            return 0.0
        return per_level * 2 * (math.ceil(self.nt / segment) - 1)

    def workspace_bytes(self, B):
        self.size_queries += 1
        o, per_shot = self.opts, 4.0 * self.nt * self._level
        shots = B * self.ns
        small = 4.0 * self._level * B * (2 * self.ns + 4)
        seg = o["history_segment"]
        if seg and seg < self.nt:                           # checkpoints in time: one recomputed segment
            return small + 4.0 * self._level * shots * (seg - 1)
        if o["adj_mode"] == 1 and not seg:                  # fused adjoint: no scratch history
            return small
        if self.clustered and o["imaging"] != 1 and not seg:  # resident imaging (default): the adjoint field stays on chip
            return small
        cap = o["scratch_mb"] * 1e6 if o["scratch_mb"] else (55 * GB if seg else 40 * GB)
        chunk = o["u_chunk_shots"] or max(1, int(cap // per_shot))
        if self.clustered and not o["u_chunk_shots"]:
            chunk = 2 * self.wave if chunk >= 2 * self.wave else (self.wave if chunk >= self.wave else chunk)
        chunk = min(chunk, shots)
        return small + chunk * per_shot * (2 if seg else 1)  # recompute tier: forward + adjoint-field scratch


def _op(monkeypatch, free_gb):
    from red_diffeq_b200 import FWIForward
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda device=None: (int(free_gb * GB), int(192 * GB)))
    monkeypatch.setattr(torch.cuda, "memory_reserved", lambda device=None: 0)
    monkeypatch.setattr(torch.cuda, "memory_allocated", lambda device=None: 0)
    ctx = dict(n_grid=70, nt=1000, dx=10.0, dt=0.001, nbc=120, f=15.0, sz=10, gz=10, ng=70, ns=5)
    return FWIForward(ctx, "cuda:0", normalize=False)      # constructing the operator loads the library, touches no device


def test_everything_fits_keeps_every_level(monkeypatch):
    op, plan = _op(monkeypatch, 178), FakePlan()
    assert op._choose_segment(plan, 64, torch.device("cuda:0")) == (0, {})  # 123.8 GB history + 24.8 GB scratch + planes
    assert plan.opts["adj_mode"] == 0 and plan.opts["u_chunk_shots"] == 0
    n = plan.size_queries
    assert op._choose_segment(plan, 64, torch.device("cuda:0")) == (0, {})  # cached: no new size queries, only the options set
    assert plan.size_queries == n and plan.log[-1] == ("history_segment", 0)
    op.set_option("img_prefetch", 8)                                        # any option change invalidates the cached decisions
    assert op._segment_auto == {}


def test_batch_too_large_for_a_history_recomputes_the_forward_field(monkeypatch):
    op, plan = _op(monkeypatch, 178), FakePlan()
    seg, _ = op._choose_segment(plan, 96, torch.device("cuda:0"))          # 185.7 GB of history: does not fit
    assert seg == plan.nt and plan.history_bytes(96, seg) == 0 and plan.opts["history_segment"] == plan.nt


def test_long_record_prefers_recomputing_over_the_fused_adjoint(monkeypatch):
    op = _op(monkeypatch, 178)
    plan = FakePlan(nt=16000, level_floats=310 * 432, wave=22)             # 8.6 GB per shot: < a wave of shots per 40 GB
    seg, _ = op._choose_segment(plan, 8, torch.device("cuda:0"))
    assert seg == plan.nt and plan.opts["adj_mode"] == 0


def test_long_record_that_fits_is_kept_unless_the_split_adjoint_is_asked_for(monkeypatch):
    """bench workload overthrust_long: 8 Marmousi-shaped models x nt = 4000 = 85.7 GB of history.  With the imaging sums formed
    in the adjoint sweep (default) nothing else is needed and every level is kept; the split adjoint's scratch history would
    hold less than a wave of shots there, so with imaging = 1 the policy recomputes the forward field instead."""
    op = _op(monkeypatch, 178)
    plan = FakePlan(nt=4000, level_floats=310 * 432, wave=22)
    assert op._choose_segment(plan, 8, torch.device("cuda:0")) == (0, {})
    op.set_option("imaging", 1)
    plan.opts["imaging"] = 1
    seg, _ = op._choose_segment(plan, 8, torch.device("cuda:0"))
    assert seg == plan.nt


def test_little_memory_falls_back_to_checkpoints_in_time(monkeypatch):
    op, plan = _op(monkeypatch, 12), FakePlan()
    seg, _ = op._choose_segment(plan, 64, torch.device("cuda:0"))
    assert seg == math.ceil(math.sqrt(2.0 * plan.nt)) and 0 < plan.history_bytes(64, seg) < 12 * GB
    assert plan.opts["adj_mode"] == 0 and plan.opts["u_chunk_shots"] == 0   # the rejected tiers' options were undone


def test_per_level_engine_sizes_its_scratch_history_from_the_free_memory(monkeypatch):
    op = _op(monkeypatch, 178)
    plan = FakePlan(ns=16, level_floats=1264 * 1264, clustered=False)      # interior 1024^2: 6.4 GB per shot
    assert op._choose_segment(plan, 1, torch.device("cuda:0"))[0] == 0
    assert plan.opts["scratch_mb"] > 40000                                  # more than the default cap: larger, fewer launches
    tight = FakePlan(ns=24, level_floats=1264 * 1264, clustered=False)     # 153 GB of history: no room for a scratch history
    assert op._choose_segment(tight, 1, torch.device("cuda:0"))[0] == 0 and tight.opts["adj_mode"] == 1


def test_explicit_segment_is_respected(monkeypatch):
    op, plan = _op(monkeypatch, 178), FakePlan()
    op.set_history_segment(16)
    assert op._choose_segment(plan, 4, torch.device("cuda:0")) == (16, {}) and plan.log[-1] == ("history_segment", 16)
    assert plan.opts["adj_mode"] == 0 and plan.opts["u_chunk_shots"] == 0 and plan.opts["scratch_mb"] == 0


def test_decisions_for_different_batches_do_not_leak_into_each_other(monkeypatch):
    """forward(B1), forward(B2), backward(B1) on one plan: every decision is applied as a whole, and the backward pass
    re-applies the decision its own forward took (the options of the last decision used to stay set)."""
    op = _op(monkeypatch, 178)
    plan = FakePlan(ns=24, level_floats=1264 * 1264, clustered=False)
    tight = op._choose_segment(plan, 1, torch.device("cuda:0"))            # no room for a scratch history: fused adjoint
    assert tight == (0, {"adj_mode": 1}) and plan.opts["adj_mode"] == 1
    plan.ns = 4                                                            # "another batch": plenty of room
    roomy = op._choose_segment(plan, 2, torch.device("cuda:0"))
    assert roomy[0] == 0 and plan.opts["adj_mode"] == 0 and plan.opts["scratch_mb"] > 40000
    op._apply_policy(plan, tight)                                          # what _solve_backward does for the first graph
    assert plan.opts["adj_mode"] == 1 and plan.opts["scratch_mb"] == 0 and plan.opts["history_segment"] == 0
