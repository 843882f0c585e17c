"""ORACLE (test infrastructure only -- never imported by the product path).

torch.optim.Adam(lr, weight_decay) + MultiStepLR exactly as configured at
VAR/pretext_VAR.py:33-37,69,72-73 and utils.py:42-46, written out in numpy
(default betas (0.9, 0.999), eps 1e-8, L2 added to the gradient, no amsgrad).
PINNED by oracle/make_golden.py against torch.optim.Adam (tests/golden/adam.npz).
"""
import numpy as np


class AdamState:
    def __init__(self, n):
        self.m = np.zeros(n, dtype=np.float32)
        self.v = np.zeros(n, dtype=np.float32)
        self.step = 0


def adam_step(p, g, st, lr, weight_decay=1e-6, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place single-tensor Adam step, float32 like torch's for-each implementation."""
    f = np.float32
    st.step += 1
    g = g.astype(np.float32) + f(weight_decay) * p
    st.m[:] = f(beta1) * st.m + f(1 - beta1) * g
    st.v[:] = f(beta2) * st.v + f(1 - beta2) * g * g
    bc1 = 1 - beta1 ** st.step
    bc2 = 1 - beta2 ** st.step
    step_size = lr / bc1
    denom = np.sqrt(st.v) / f(np.sqrt(bc2)) + f(eps)
    p -= f(step_size) * (st.m / denom)
    return p


def multistep_lr(base_lr, epoch, milestones, gamma):
    """LR in effect during epoch `epoch` (scheduler.step() called once per finished epoch)."""
    k = sum(1 for m in milestones if epoch >= m)
    return base_lr * (gamma ** k)
