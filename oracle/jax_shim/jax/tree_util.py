"""Minimal pytree utilities (dict with sorted keys, list, tuple, None, registered classes)."""
import functools

_REGISTRY = {}


class GetAttrKey:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f".{self.name}"


class DictKey:
    def __init__(self, key):
        self.key = key

    def __repr__(self):
        return f"[{self.key!r}]"


class SequenceKey:
    def __init__(self, idx):
        self.idx = idx

    def __repr__(self):
        return f"[{self.idx}]"


def register_pytree_with_keys(cls, flatten_with_keys, unflatten, flatten_func=None):
    _REGISTRY[cls] = (flatten_with_keys, unflatten)


def register_pytree_node(cls, flatten, unflatten):
    def with_keys(x):
        children, aux = flatten(x)
        return tuple((SequenceKey(i), c) for i, c in enumerate(children)), aux

    _REGISTRY[cls] = (with_keys, unflatten)


class Partial(functools.partial):
    pass


def _children(x):
    """Returns (kind, keys, children, rebuild) or None for a leaf."""
    if x is None:
        return ("none", (), (), lambda ch: None)
    if isinstance(x, dict):
        keys = sorted(x.keys())
        return ("dict", tuple(DictKey(k) for k in keys), tuple(x[k] for k in keys),
                lambda ch, keys=keys: {k: c for k, c in zip(keys, ch)})
    if isinstance(x, list):
        return ("list", tuple(SequenceKey(i) for i in range(len(x))), tuple(x), lambda ch: list(ch))
    if isinstance(x, tuple):
        return ("tuple", tuple(SequenceKey(i) for i in range(len(x))), tuple(x), lambda ch: tuple(ch))
    for cls, (flat, unflat) in _REGISTRY.items():
        if type(x) is cls:
            kc, aux = flat(x)
            keys = tuple(k for k, _ in kc)
            ch = tuple(c for _, c in kc)
            return ("node:" + cls.__name__, keys, ch, lambda ch, aux=aux, unflat=unflat: unflat(aux, ch))
    return None


def tree_map(f, tree, *rest, is_leaf=None):
    if is_leaf is not None and is_leaf(tree):
        return f(tree, *rest)
    node = _children(tree)
    if node is None:
        return f(tree, *rest)
    kind, keys, ch, rebuild = node
    rest_ch = []
    for r in rest:
        rn = _children(r)
        if rn is None or len(rn[2]) != len(ch):
            raise ValueError("tree_map: structure mismatch")
        rest_ch.append(rn[2])
    return rebuild([tree_map(f, c, *[rc[i] for rc in rest_ch], is_leaf=is_leaf) for i, c in enumerate(ch)])


def tree_leaves(tree, is_leaf=None):
    out = []

    def rec(x):
        if is_leaf is not None and is_leaf(x):
            out.append(x)
            return
        node = _children(x)
        if node is None:
            out.append(x)
            return
        for c in node[2]:
            rec(c)

    rec(tree)
    return out


def tree_structure(tree, is_leaf=None):
    def rec(x):
        if is_leaf is not None and is_leaf(x):
            return "*"
        node = _children(x)
        if node is None:
            return "*"
        kind, keys, ch, _ = node
        return (kind, tuple(repr(k) for k in keys), tuple(rec(c) for c in ch))

    return rec(tree)


def tree_flatten(tree, is_leaf=None):
    return tree_leaves(tree, is_leaf), tree_structure(tree, is_leaf)


def tree_flatten_with_path(tree, is_leaf=None):
    out = []

    def rec(x, path):
        if is_leaf is not None and is_leaf(x):
            out.append((path, x))
            return
        node = _children(x)
        if node is None:
            out.append((path, x))
            return
        for k, c in zip(node[1], node[2]):
            rec(c, path + (k,))

    rec(tree, ())
    return out, tree_structure(tree, is_leaf)
