"""ORACLE (test infrastructure only; imported by tests/, never by the package): the optimiser step of the reference
training loop restated with torch ops in float64 - `torch.nn.utils.clip_grad_norm_(params, max_norm)` (train.py:227)
followed by the single-tensor AdamW of torch/optim/adamw.py that `torch.optim.AdamW` (train.py:79-83) runs.  Pinned by
`tests/test_optim.py::test_reference_step_is_clip_then_adamw` against PyTorch's own clip + AdamW on the CPU."""
import math

import torch  # noqa: F401


def reference_step(params, grads, exp_avg, exp_avg_sq, step, lr, betas, eps, weight_decay, max_grad_norm):
    """One clip + AdamW step in float64; returns ([(param, exp_avg, exp_avg_sq)], pre-clip global norm)."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    coef = 1.0 if max_grad_norm is None else min(1.0, max_grad_norm / (total + 1e-6))
    b1, b2 = betas
    out = []
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        p, g, m, v = p.double(), g.double() * coef, m.double(), v.double()
        p = p * (1 - lr * weight_decay)
        m = m + (1 - b1) * (g - m)
        v = b2 * v + (1 - b2) * g * g
        denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
        p = p - (lr / (1 - b1 ** step)) * m / denom
        out.append((p, m, v))
    return out, total
