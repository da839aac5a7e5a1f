"""Oracle restatement of `LREQAdam.step` (reference model/utils/custom_adam.py:24-76).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import math

import torch


def lreq_adam_step(params, grads, exp_avg_sq, steps, coefs, lr, beta2=0.99, eps=1e-8):
    """In-place on `params` / `exp_avg_sq` (lists of fp32 tensors); `steps` = per-tensor step counters (already
    incremented), `coefs` = per-tensor lr_equalization_coef (None = attribute absent); grads[i] None = skipped."""
    for p, g, v, t, c in zip(params, grads, exp_avg_sq, steps, coefs):
        if g is None:
            continue
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)                # :62
        denom = v.sqrt().add_(eps)                                   # :63
        step_size = lr * math.sqrt(1 - beta2 ** t)                   # :66-68
        if c is not None:
            step_size *= c                                           # :71-72
        p.addcdiv_(g, denom, value=-step_size)                       # :74
