"""CPU: SemiStep's event-bracket helpers (`_event`, `_phase`) -- inert without a `timers` dict, selective with one --
with torch.cuda.Event replaced by a counter (the step itself needs a GPU)."""
import torch

from u2pl_b200.step import SemiStep


class _Ev:
    n = 0

    def __init__(self, enable_timing=True):
        type(self).n += 1
        self.id = type(self).n

    def record(self):
        pass


def test_phase_brackets(monkeypatch):
    monkeypatch.setattr(torch.cuda, "Event", _Ev)
    s = SemiStep.__new__(SemiStep)
    assert s._event() is None and s._phase(None, "t1") is None             # no timers installed: nothing is recorded
    s.timers = {}
    assert s._event() is None
    s.timers = {"entropy_partition": [], "t2": []}
    a = s._event()
    assert a is not None
    b = s._phase(a, "t1")                                                  # "t1" is not being timed: closed silently
    assert s.timers["t2"] == [] and b is not None
    c = s._phase(b, "t2")
    assert len(s.timers["t2"]) == 1 and s.timers["t2"][0][0] is b and c is not None
    assert set(SemiStep.PHASES) >= {"t1", "t2", "backward"}
