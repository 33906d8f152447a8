"""Train state and a minimal optimiser (mirror of gpfy/training.py:24-99; optax is not installable, so
`adam` / `sgd` follow optax's (init, update) protocol on nested dicts of numpy arrays).  O(#params) host
work per step — the caller of the hot path, not part of it."""
import dataclasses
from typing import Any, Callable, Dict

import numpy as np

from .param import tree_map


@dataclasses.dataclass(frozen=True)
class GradientTransformation:
    init: Callable
    update: Callable


def adam(learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> GradientTransformation:
    def init(params):
        z = tree_map(lambda p: np.zeros_like(p), params)
        return {"count": 0, "mu": z, "nu": tree_map(lambda p: np.zeros_like(p), params)}

    def update(grads, state, params=None):
        c = state["count"] + 1
        mu = tree_map(lambda m, g: b1 * m + (1 - b1) * g, state["mu"], grads)
        nu = tree_map(lambda v, g: b2 * v + (1 - b2) * g * g, state["nu"], grads)
        upd = tree_map(lambda m, v: -learning_rate * (m / (1 - b1**c)) / (np.sqrt(v / (1 - b2**c)) + eps), mu, nu)
        return upd, {"count": c, "mu": mu, "nu": nu}

    return GradientTransformation(init, update)


def sgd(learning_rate: float) -> GradientTransformation:
    return GradientTransformation(lambda params: {}, lambda g, s, p=None: (tree_map(lambda x: -learning_rate * x, g), s))


def masked(tx: GradientTransformation, trainables) -> GradientTransformation:
    """optax.chain(tx, optax.masked(optax.set_to_zero(), frozen)) of model.py:288-289."""
    def update(grads, state, params=None):
        upd, st = tx.update(grads, state, params)
        return tree_map(lambda u, t: u if t else np.zeros_like(u), upd, trainables), st
    return GradientTransformation(tx.init, update)


@dataclasses.dataclass(frozen=True)
class TrainState:
    step: int
    apply_fn: Callable
    params: Dict[str, Any]
    tx: GradientTransformation
    opt_state: Any

    def apply_gradients(self, *, grads, **kwargs):
        updates, new_opt_state = self.tx.update(grads, self.opt_state, self.params)
        new_params = tree_map(lambda p, u: p + u, self.params, updates)
        return dataclasses.replace(self, step=self.step + 1, params=new_params, opt_state=new_opt_state, **kwargs)

    @classmethod
    def create(cls, *, apply_fn, params, tx, **kwargs):
        return cls(step=0, apply_fn=apply_fn, params=params, tx=tx, opt_state=tx.init(params), **kwargs)
