"""`Param`: the explicit parameter container of the gpfy API (mirror of gpfy/param.py:28-258).

Same attributes and methods as the reference (`params`, `_trainables`, `_bijectors`, `constants`,
`_constrained`, `replace`, `replace_param`, `set_trainable`, `set_bijector`, `constrained`,
`unconstrained`) over plain nested dicts of float64 numpy arrays; the bijectors are numpy
restatements of the three TFP bijectors the reference uses (Exp, Identity, FillTriangular).
"""
import copy
import dataclasses
import math
from typing import Any, Dict

import numpy as np


class Bijector:
    def forward(self, x):
        raise NotImplementedError

    def inverse(self, y):
        raise NotImplementedError

    def forward_vjp(self, x, g):
        """Cotangent w.r.t. the unconstrained x given cotangent g w.r.t. forward(x)."""
        raise NotImplementedError


class Exp(Bijector):  # param.py:24 `positive`
    def forward(self, x):
        return np.exp(x)

    def inverse(self, y):
        return np.log(y)

    def forward_vjp(self, x, g):
        return g * np.exp(x)


class Identity(Bijector):  # param.py:25
    def forward(self, x):
        return x

    def inverse(self, y):
        return y

    def forward_vjp(self, x, g):
        return g


class FillTriangular(Bijector):
    """TFP FillTriangular(upper=False) (variational.py:260): vector n(n+1)/2 -> lower [n, n] by
    concat([x[n:], reverse(x)]) reshaped [n, n] and masked to the lower band."""

    _maps = {}

    @classmethod
    def _index_map(cls, n):
        """(flat position in the [n, n] matrix, position in the vector) of every lower-triangular entry, cached."""
        if n not in cls._maps:
            m = n * (n + 1) // 2
            x = np.arange(1, m + 1)
            xc = np.tril(np.concatenate([x[n:], x[::-1]]).reshape(n, n))
            ii, jj = np.nonzero(xc)
            cls._maps[n] = (ii * n + jj, xc[ii, jj] - 1)
        return cls._maps[n]

    def forward(self, x):
        x = np.asarray(x)
        m = x.shape[-1]
        n = int(round((math.sqrt(1 + 8 * m) - 1) / 2))
        flat, src = self._index_map(n)
        out = np.zeros(x.shape[:-1] + (n * n,), dtype=x.dtype)
        out[..., flat] = x[..., src]
        return out.reshape(x.shape[:-1] + (n, n))

    def inverse(self, y):
        y = np.asarray(y)
        n = y.shape[-1]
        flat, src = self._index_map(n)
        out = np.zeros(y.shape[:-2] + (n * (n + 1) // 2,), dtype=y.dtype)
        out[..., src] = y.reshape(y.shape[:-2] + (n * n,))[..., flat]
        return out

    def forward_vjp(self, x, g):
        return self.inverse(g)  # gather of the lower triangle


positive = Exp
identity = Identity


def tree_map(f, t, *rest):
    if isinstance(t, dict):
        return {k: tree_map(f, t[k], *[r[k] for r in rest]) for k in t}
    return f(t, *rest)


def tree_leaves(t):
    if isinstance(t, dict):
        out = []
        for k in sorted(t):
            out.extend(tree_leaves(t[k]))
        return out
    if isinstance(t, (list, tuple)):
        out = []
        for v in t:
            out.extend(tree_leaves(v))
        return out
    return [t]


def _structure(t):
    if isinstance(t, dict):
        return tuple((k, _structure(t[k])) for k in sorted(t))
    return "*"


@dataclasses.dataclass(frozen=True)
class Param:
    params: Dict[str, Any] = dataclasses.field(default_factory=dict)
    _trainables: Dict[str, Any] = dataclasses.field(default_factory=dict)
    _bijectors: Dict[str, Any] = dataclasses.field(default_factory=dict)
    constants: Dict[str, Any] = dataclasses.field(default_factory=dict)
    _constrained: bool = True

    # ---- validation (param.py:53-103) ----
    def _is_subtree(self, t1, t2) -> bool:
        if isinstance(t1, dict) and isinstance(t2, dict):
            return all((k in t2) and self._is_subtree(t1[k], t2[k]) for k in t1)
        if isinstance(t2, dict):
            return False
        return True

    def _tree_update_from_subtree(self, t1, t2):
        ret = {}
        for k1, v1 in t1.items():
            if k1 in t2:
                ret[k1] = self._tree_update_from_subtree(t1[k1], t2[k1]) if isinstance(v1, dict) else t2[k1]
            else:
                ret[k1] = self._tree_update_from_subtree(t1[k1], t2) if isinstance(v1, dict) else v1
        return ret

    def __post_init__(self):
        if not self._is_subtree(self._trainables, self.params):
            raise ValueError("Invalid key in `_trainables`")
        if not self._is_subtree(self._bijectors, self.params):
            raise ValueError("Invalid key in `_bijectors`")
        if not all(c in self.params or c == "sphere" for c in self.constants):
            raise ValueError("Invalid key in `_constants`")
        trainables = self._trainables
        if not trainables or _structure(trainables) != _structure(self.params):
            trainables = self._tree_update_from_subtree(tree_map(lambda _: True, self.params), self._trainables)
        bijectors = self._bijectors
        if not bijectors or _structure(bijectors) != _structure(self.params):
            bijectors = self._tree_update_from_subtree(tree_map(lambda _: positive(), self.params), self._bijectors)
        params = tree_map(lambda x: np.array(x, dtype=np.float64), self.params)  # param.py:151
        object.__setattr__(self, "params", params)
        object.__setattr__(self, "_trainables", trainables)
        object.__setattr__(self, "_bijectors", bijectors)

    def replace(self, **updates) -> "Param":
        return dataclasses.replace(self, **updates)

    def replace_param(self, collection: str, **kwargs) -> "Param":
        if collection not in self.params:
            raise ValueError(f"there is no {collection} collection")
        updates = self._tree_update_from_subtree(self.params[collection], kwargs)
        updates = self._tree_update_from_subtree(self.params, {collection: updates})
        return self.replace(params=updates)

    def set_trainable(self, collection: str, **kwargs) -> "Param":
        if collection not in self._trainables:
            raise ValueError(f"there is no {collection} collection")
        updates = self._tree_update_from_subtree(self._trainables[collection], kwargs)
        updates = self._tree_update_from_subtree(self._trainables, {collection: updates})
        return self.replace(_trainables=updates)

    def set_bijector(self, collection: str, **kwargs) -> "Param":
        if collection not in self._trainables:
            raise ValueError(f"there is no {collection} collection")
        updates = self._tree_update_from_subtree(self._bijectors[collection], kwargs)
        updates = self._tree_update_from_subtree(self._bijectors, {collection: updates})
        return self.replace(_bijectors=updates)

    def unconstrained(self) -> "Param":
        free = tree_map(lambda p, t: t.inverse(p), self.params, self._bijectors)
        return self.replace(_constrained=False, params=free)

    def constrained(self) -> "Param":
        con = tree_map(lambda p, t: t.forward(p), self.params, self._bijectors)
        return self.replace(_constrained=True, params=con)
