"""CPU: LRSchedule vs the reference's own class (imported from /root/reference when present, else the committed trace
in tests/golden/lr_schedule.json generated from it)."""
import importlib.util
import json
import os

from latex_ocr_b200.lr_schedule import LRSchedule

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lr_schedule.json")
REF = "/root/reference/model/utils/lr_schedule.py"
CASES = [
    dict(lr_init=1e-3, lr_min=1e-5, start_decay=6, end_decay=13, lr_warm=1e-4, end_warm=2),          # train.py:49-56 shape
    dict(lr_init=1e-3, lr_min=1e-4, start_decay=0, decay_rate=0.5, early_stopping=3),
    dict(lr_init=2e-3, lr_min=1e-4, start_decay=3, end_decay=9),
    dict(lr_init=1e-3),
]
SCORES = [None, 1.0, 0.5, None, 0.5, 0.7, 0.6, None, 0.6, 0.5, 0.4, None, None, 0.3, None, None]


def _trace(cls, kw):
    s = cls(**kw)
    out = []
    for i, sc in enumerate(SCORES):
        s.update(batch_no=i)
        if sc is not None:
            s.update(score=sc)
        out.append((s.lr, bool(s.stop_training)))
    return out


def _reference_traces():
    if os.path.exists(REF):
        spec = importlib.util.spec_from_file_location("ref_lr_schedule", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        tr = [_trace(mod.LRSchedule, kw) for kw in CASES]
        with open(GOLD, "w") as f:
            json.dump(tr, f)
        return tr
    return [[tuple(x) for x in t] for t in json.load(open(GOLD))]


def test_lr_schedule_matches_reference():
    for kw, want in zip(CASES, _reference_traces()):
        got = _trace(LRSchedule, kw)
        for (lr, stop), (wlr, wstop) in zip(got, want):
            assert abs(lr - wlr) <= 1e-15 + 1e-12 * abs(wlr) and stop == bool(wstop), (kw, got, want)
