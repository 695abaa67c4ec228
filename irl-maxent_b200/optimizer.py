"""
optimizer -- gradient-ASCENT steppers, learning-rate schedules and parameter
initialisers with the interface of the reference's `optimizer.py`
(`/root/reference/src/optimizer.py`), working on numpy arrays AND on
device-resident torch tensors.

The IRL loops keep omega on the GPU and expect `step` to update it *in place*
through the object handed to `reset` (reference: optimizer.py:33,107,164 --
`maxent.irl` never reads a return value).  All updates below are therefore
in-place and type-agnostic: `_exp`, `_norm` dispatch on the array type, the
arithmetic is otherwise plain operator syntax that numpy and torch share.
"""

import numpy as np


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


def _exp(x):
    if _is_tensor(x):
        import torch
        return torch.exp(x)
    return np.exp(x)


def _norm(x, ord=None):
    if _is_tensor(x):
        import torch
        return torch.linalg.vector_norm(x, ord=2 if ord is None else ord)
    return np.linalg.norm(x, ord)


def _rate(lr, k):
    """A schedule `(k) -> lr` or a constant (reference: optimizer.py:104,161)."""
    # plain Python float: a numpy scalar times a CUDA tensor would try to pull the tensor to the host
    return float(lr(k) if callable(lr) else lr)


class Optimizer:
    """Base class: remembers (does not copy) the parameter array (optimizer.py:12-58)."""

    def __init__(self):
        self.parameters = None

    def reset(self, parameters):
        self.parameters = parameters

    def step(self, grad, *args, **kwargs):
        raise NotImplementedError

    def normalize_grad(self, ord=None):
        """Wrap this optimizer so that it sees grad / ||grad||_ord."""
        return NormalizeGrad(self, ord)


class _Counted(Optimizer):
    """Steppers that count their steps for the schedule (k starts at 0 after reset)."""

    def __init__(self, lr):
        super().__init__()
        self.lr = lr
        self.k = 0

    def reset(self, parameters):
        super().reset(parameters)
        self.k = 0

    def _next_rate(self):
        lr = _rate(self.lr, self.k)
        self.k += 1
        return lr


class Sga(_Counted):
    """theta += lr_k * grad (reference: optimizer.py:61-107)."""

    def step(self, grad, *args, **kwargs):
        lr = self._next_rate()
        self.parameters += lr * grad


class ExpSga(_Counted):
    """theta *= exp(lr_k * grad), optionally renormalised to sum 1
    (reference: optimizer.py:110-167; Ziebart 2010 Alg. 10.5)."""

    def __init__(self, lr, normalize=False):
        super().__init__(lr)
        self.normalize = normalize

    def step(self, grad, *args, **kwargs):
        lr = self._next_rate()
        self.parameters *= _exp(lr * grad)
        if self.normalize:
            self.parameters /= self.parameters.sum()


class NormalizeGrad(Optimizer):
    """Delegates to `opt` with the gradient scaled to unit norm
    (reference: optimizer.py:170-214; `ord` as in numpy.linalg.norm)."""

    def __init__(self, opt, ord=None):
        super().__init__()
        self.opt = opt
        self.ord = ord

    def reset(self, parameters):
        super().reset(parameters)
        self.opt.reset(parameters)

    def step(self, grad, *args, **kwargs):
        return self.opt.step(grad / _norm(grad, self.ord), *args, **kwargs)


# -- schedules (reference: optimizer.py:217-293) ----------------------------------------

def linear_decay(lr0=0.2, decay_rate=1.0, decay_steps=1):
    """lr0 / (1 + decay_rate * floor(k / decay_steps))"""
    return lambda k: lr0 / (1.0 + decay_rate * np.floor(k / decay_steps))


def power_decay(lr0=0.2, decay_rate=1.0, decay_steps=1, power=2):
    """lr0 / (1 + decay_rate * floor(k / decay_steps)) ** power"""
    return lambda k: lr0 / (decay_rate * np.floor(k / decay_steps) + 1.0) ** power


def exponential_decay(lr0=0.2, decay_rate=0.5, decay_steps=1):
    """lr0 * exp(-decay_rate * floor(k / decay_steps))"""
    return lambda k: lr0 * np.exp(-decay_rate * np.floor(k / decay_steps))


# -- initialisers (reference: optimizer.py:296-398) ---------------------------------------

class Initializer:
    def initialize(self, shape):
        raise NotImplementedError

    def __call__(self, shape):
        return self.initialize(shape)


class Uniform(Initializer):
    """U[low, high) from numpy's global generator (same stream as the reference, :366)."""

    def __init__(self, low=0.0, high=1.0):
        self.low, self.high = low, high

    def initialize(self, shape):
        return np.random.uniform(size=shape, low=self.low, high=self.high)


class Constant(Initializer):
    """A constant, or a function of the shape returning one (reference: :369-398)."""

    def __init__(self, value=1.0):
        self.value = value

    def initialize(self, shape):
        v = self.value(shape) if callable(self.value) else self.value
        return np.ones(shape) * v
