"""Opt-in CUDA-graph replay of the fused training nodes whose parameters are FROZEN (the synthesis pass, train_g.py; the
LPIPS-VGG16 distance, train_lpips.py).  SURVEY 8f-3: at batch 1 (the inversion loop, embedding_img.py:84-128) an iteration
is bound by the host's launch calls, not by the GPU.

Switch: `DGE_TRAIN_GRAPHS=1` in the environment, or `dge_b200.graphs.GRAPHS = True`.

After `WARMUP` eager passes of one configuration (owner module, slot = input shapes / flags, key = weights epoch) a node's
forward chain is captured into a CUDA graph over a private memory pool, and the first backward after that into a second
graph over the same pool.  Later passes copy their inputs into the graphs' static inputs, replay, and hand out CLONES of
the static outputs, so nothing the caller holds aliases graph memory.  The activations a backward reads live in the pool
and are overwritten by the next forward replay of the same slot: a backward through an OLDER pass than the slot's latest
(two passes of one shape alive, backward through the first) cannot be served and raises -- that pattern needs the switch
off.  `retain_graph=True` + a second backward of the latest pass (E_align_s2.py:205,220; embedding_img.py:100-128) replays
the same backward graph.  Any failure while capturing disables the graphs of that slot with a warning: the eager chain is
always the fallback, never a different result.
"""
import collections
import os
import warnings

import torch

GRAPHS = os.environ.get('DGE_TRAIN_GRAPHS', '0') == '1'
WARMUP = 2
MAX_SLOTS = 8       # replay states kept per owner module, least recently used first out: scripts that crop to a new random
                    # size every iteration (E_align_cropping_s1.py:150-170) would otherwise collect a memory pool per size


class State:
    def __init__(self, key):
        self.key, self.calls, self.gen = key, 0, 0
        self.failed = False
        self.fwd = self.bwd = None          # torch.cuda.CUDAGraph
        self.ins = self.outs = self.saved = None
        self.bwd_ins = self.bwd_outs = None


def state_for(owner, slot, key):
    """The replay state of (owner module, slot); a changed key (new weights) starts over and drops the old pool."""
    table = owner.__dict__.setdefault('_dge_graphs', collections.OrderedDict())
    st = table.get(slot)
    if st is None or st.key != key:
        st = table[slot] = State(key)
    table.move_to_end(slot)
    while len(table) > MAX_SLOTS:            # a dropped state lives on while a pass that used it still holds its handle
        table.popitem(last=False)
    return st


def stale(what):
    return RuntimeError(
        f'dge_b200 {what}: backward through a pass whose saved activations were overwritten by a later pass of the same '
        'shape (CUDA-graph mode keeps ONE pass per shape alive); set dge_b200.graphs.GRAPHS = False / unset '
        'DGE_TRAIN_GRAPHS for this pattern')


def capture(fn, pool=None):
    """Capture fn() into a new CUDA graph -> (graph, what fn returned: tensors in the graph's pool)."""
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, pool=pool):
        out = fn()
    return g, out


def failed(st, what, exc):
    warnings.warn(f'dge_b200 {what}: CUDA-graph capture failed ({exc!r}); running eagerly')
    st.failed = True
    torch.cuda.synchronize()


def forward(owner, slot, key, inputs, fwd, what, enabled=True):
    """Run a node's forward: fwd(*inputs) -> (outs: tuple of tensors, saved: anything the backward reads).
    -> (outs, handle for `backward`).  Eager unless the switch is on, `enabled`, and the slot has warmed up."""
    st = None
    if GRAPHS and enabled:
        st = state_for(owner, slot, key)
        if st.failed:
            st = None
    if st is not None and st.fwd is None and st.calls >= WARMUP:
        try:
            st.ins = tuple(t.detach().contiguous().clone() for t in inputs)
            st.fwd, (st.outs, st.saved) = capture(lambda: fwd(*st.ins))
        except Exception as exc:  # noqa: BLE001 -- the eager chain is the fallback
            failed(st, what + ' forward', exc)
            st.fwd = None
            st = None
    if st is not None and st.fwd is not None:
        for s, t in zip(st.ins, inputs):
            s.copy_(t)
        st.fwd.replay()
        st.gen += 1
        return tuple(o.clone() for o in st.outs), (st, st.gen, None)
    if st is not None:
        st.calls += 1
    outs, saved = fwd(*inputs)
    return outs, (None, 0, saved)


def backward(handle, grads, bwd, what):
    """Run the node's backward: bwd(saved, *grads) -> a tensor or a tuple of tensors / None."""
    st, gen, saved = handle
    if st is None:
        return bwd(saved, *grads)
    if gen != st.gen:
        raise stale(what)
    if st.failed:
        return bwd(st.saved, *grads)
    if st.bwd is None:
        try:
            st.bwd_ins = tuple(g.detach().contiguous().clone() for g in grads)
            st.bwd, st.bwd_outs = capture(lambda: bwd(st.saved, *st.bwd_ins), pool=st.fwd.pool())
        except Exception as exc:  # noqa: BLE001 -- this pass's activations are intact: finish it eagerly
            failed(st, what + ' backward', exc)
            return bwd(st.saved, *grads)
    for s, g in zip(st.bwd_ins, grads):
        s.copy_(g)
    st.bwd.replay()
    outs = st.bwd_outs
    if isinstance(outs, tuple):
        return tuple(None if o is None else o.clone() for o in outs)
    return outs.clone()
