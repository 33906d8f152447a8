"""Eager control flow."""
import numpy as _np

from . import tree_util as _tu


_UNSET = object()


def cond(pred, true_fun, false_fun, *operands, operand=_UNSET):
    if operand is not _UNSET and not operands:
        operands = (operand,)
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def scan(f, init, xs=None, length=None):
    carry = init
    ys = []
    if xs is None:
        for _ in range(int(length)):
            carry, y = f(carry, None)
            ys.append(y)
    else:
        n = _np.shape(_tu.tree_leaves(xs)[0])[0]
        for i in range(n):
            carry, y = f(carry, _tu.tree_map(lambda a: a[i], xs))
            ys.append(y)
    if ys and ys[0] is not None:
        ys = _tu.tree_map(lambda *v: _np.stack(v, 0), *ys)
    else:
        ys = None
    return carry, ys


def switch(index, branches, *operands):
    return branches[int(index)](*operands)
