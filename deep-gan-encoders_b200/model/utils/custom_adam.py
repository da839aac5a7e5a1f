"""LREQAdam -- drop-in for the reference's `model/utils/custom_adam.py` (:6-76).

Adam with beta1 == 0 (no first moment) whose step is scaled per parameter by the
`lr_equalization_coef` attribute the `ln.*` layers attach (lreq.py:60-62, 118-120).  Same constructor
checks, `state` layout (`step`, `exp_avg_sq`) and update rule; the per-parameter Python loop of five
tiny kernels becomes ONE multi-tensor kernel launch (`dge_lreq_adam_step`).
"""
import ctypes
import math

import torch
from torch.optim.optimizer import Optimizer

from dge_b200 import ops

_CHUNK = 65536


class LREQAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.0, 0.99), eps=1e-8, weight_decay=0):
        beta_2 = betas[1]
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 == betas[0]:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= beta_2 < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(beta_2))
        defaults = dict(lr=lr, beta_2=beta_2, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._plan_key, self._plan = None, None
        # Data-parallel runs (one process per GPU under torchrun, torch.distributed initialised): the reference loop
        # (E_align_s2.py:203-206, 218-221) is `zero_grad(); loss.backward(); step()` and knows nothing about ranks, so
        # the gradient exchange lives here -- the parameters' .grad become views of one flat bucket whose all-reduce
        # overlaps the backward (dge_b200.dist.GradBucket) and `step` waits for it before the update.
        self._bucket = None
        from dge_b200 import dist as _ddist
        if _ddist.world_size() > 1:
            ps = [p for g in self.param_groups for p in g['params']]
            if ps and all(p.is_cuda or _ddist.dist.get_backend() == 'gloo' for p in ps):
                self._bucket = _ddist.GradBucket(ps)

    def zero_grad(self, set_to_none=True):
        if self._bucket is not None:
            self._bucket.zero()
            return
        super().zero_grad(set_to_none)

    def _build_plan(self, tensors, device):
        """Device-side tables for the multi-tensor launch; rebuilt only when the (p, grad, v) pointers change."""
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]['exp_avg_sq'].data_ptr(), p.numel()) for p in tensors)
        if key == self._plan_key:
            return self._plan
        i64 = lambda v: torch.tensor(v, dtype=torch.int64, device=device)
        blk_t, blk_o = [], []
        for t, p in enumerate(tensors):
            for off in range(0, p.numel(), _CHUNK):
                blk_t.append(t)
                blk_o.append(off)
        plan = dict(params=i64([k[0] for k in key]), grads=i64([k[1] for k in key]), vs=i64([k[2] for k in key]),
                    numel=i64([k[3] for k in key]), blk_t=torch.tensor(blk_t, dtype=torch.int32, device=device),
                    blk_o=i64(blk_o), n_blocks=len(blk_t))
        self._plan_key, self._plan = key, plan
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._bucket is not None:
            self._bucket.finish()
        for group in self.param_groups:
            tensors, steps = [], []
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients, please consider SparseAdam instead')
                if not p.is_cuda:
                    raise ops.DgeError('LREQAdam: dge_b200 runs on a B200 only; parameters must be CUDA tensors')
                if group['weight_decay'] != 0:
                    # the reference dereferences the non-existent attribute `p.coef` here (custom_adam.py:57)
                    raise AttributeError("'Parameter' object has no attribute 'coef'")
                if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise ops.DgeError('LREQAdam: contiguous fp32 parameters / gradients expected')
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg_sq'] = torch.zeros_like(p.data)
                state['step'] += 1
                step_size = group['lr'] * math.sqrt(1 - group['beta_2'] ** state['step'])
                if hasattr(p, 'lr_equalization_coef'):
                    step_size *= p.lr_equalization_coef
                tensors.append(p)
                steps.append(step_size)
            if not tensors:
                continue
            dev = tensors[0].device
            plan = self._build_plan(tensors, dev)
            step_t = torch.tensor(steps, dtype=torch.float32, device=dev)
            vp = lambda t: ctypes.c_void_p(t.data_ptr())
            ops.check(ops.lib().dge_lreq_adam_step(vp(plan['params']), vp(plan['grads']), vp(plan['vs']),
                                                   vp(plan['numel']), vp(step_t), vp(plan['blk_t']), vp(plan['blk_o']),
                                                   plan['n_blocks'], _CHUNK, float(group['beta_2']),
                                                   float(group['eps']), ops._stream()))
        # the kernel writes the parameters through raw pointers (no torch version bump): retire every tensor derived
        # from them (packed conv weights etc.) -- see ops.weight_key
        ops.invalidate_weight_caches(p for g in self.param_groups for p in g['params'])
        return loss
