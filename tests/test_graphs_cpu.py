"""Host logic of the opt-in CUDA-graph replay (dge_b200/graphs.py) that needs no device: the switch is off by default, an
eager pass goes straight through, the per-owner slot table is bounded and restarts a slot when its key (weights) changes,
and the refusal of a stale backward names the switch.  The replay itself is covered on the B200
(tests/test_train_fused_gpu.py::test_*_cuda_graph_replay_*)."""
import os
import subprocess
import sys

import torch

from dge_b200 import graphs

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-gan-encoders_b200")


class _Owner:
    pass


def test_switch_is_off_by_default_and_follows_the_environment():
    assert graphs.GRAPHS is False
    code = "from dge_b200 import graphs; print(graphs.GRAPHS)"
    env = dict(os.environ, DGE_TRAIN_GRAPHS="1", PYTHONPATH=PKG)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.stdout.strip() == "True", out.stderr[-400:]


def test_eager_pass_goes_straight_through():
    calls = []

    def fwd(a, b):
        calls.append("f")
        return (a + b, a * b), {"a": a, "b": b}

    def bwd(saved, g0, g1):
        calls.append("b")
        return g0 + g1 * saved["b"], None

    a, b = torch.arange(3.0), torch.ones(3) * 2
    outs, handle = graphs.forward(_Owner(), "slot", None, (a, b), fwd, "test")
    assert handle[0] is None and torch.equal(outs[0], a + b)
    g = graphs.backward(handle, (torch.ones(3), torch.ones(3)), bwd, "test")
    assert torch.equal(g[0], torch.full((3,), 3.0)) and g[1] is None and calls == ["f", "b"]


def test_slot_table_is_bounded_and_restarts_on_a_new_key():
    o = _Owner()
    a = graphs.state_for(o, "s0", ("w", 1))
    assert graphs.state_for(o, "s0", ("w", 1)) is a
    a.calls = 5
    b = graphs.state_for(o, "s0", ("w", 2))                  # the weights changed: warm up and capture again
    assert b is not a and b.calls == 0
    for i in range(graphs.MAX_SLOTS + 3):
        graphs.state_for(o, ("shape", i), 0)
    table = o.__dict__["_dge_graphs"]
    assert len(table) == graphs.MAX_SLOTS and "s0" not in table and ("shape", graphs.MAX_SLOTS + 2) in table
    graphs.state_for(o, ("shape", 4), 0)                      # touching a slot makes it the most recent one
    graphs.state_for(o, ("shape", 99), 0)
    assert ("shape", 4) in table and ("shape", 3) not in table


def test_stale_backward_is_refused_with_the_switch_named():
    st = graphs.State(0)
    st.gen = 3
    try:
        graphs.backward((st, 2, None), (torch.ones(1),), lambda saved, g: g, "train_x")
    except RuntimeError as exc:
        assert "DGE_TRAIN_GRAPHS" in str(exc) and "train_x" in str(exc)
    else:
        raise AssertionError("a backward through an overwritten pass must raise")
