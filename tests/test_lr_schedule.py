"""CPU: LRSchedule vs the trace of the reference's own class committed in tests/golden/lr_schedule.json (written by
oracle/make_golden_host.py, never by a test); with the reference tree present the live class must reproduce the file."""
import importlib.util
import json
import os

import pytest

from latex_ocr_b200.lr_schedule import LRSchedule
from oracle import make_golden_host as mgh
from oracle import ref_shim

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lr_schedule.json")


def _close(got, want):
    for (lr, stop), (wlr, wstop) in zip(got, want):
        assert abs(lr - wlr) <= 1e-15 + 1e-12 * abs(wlr) and stop == bool(wstop), (got, want)


def test_lr_schedule_matches_reference():
    for kw, want in zip(mgh.LR_CASES, json.load(open(GOLD))):
        _close(mgh.lr_trace(LRSchedule, kw), want)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not in this environment")
def test_committed_trace_equals_live_reference():
    spec = importlib.util.spec_from_file_location("ref_lr_schedule", os.path.join(ref_shim.REFERENCE_ROOT, "model", "utils", "lr_schedule.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for kw, want in zip(mgh.LR_CASES, json.load(open(GOLD))):
        _close(mgh.lr_trace(mod.LRSchedule, kw), want)
