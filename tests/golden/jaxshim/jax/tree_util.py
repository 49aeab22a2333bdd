"""jax.tree_util stand-in: lists, tuples, dicts (sorted keys), None and registered dataclasses are nodes."""
import dataclasses

_REGISTRY = {}   # type -> (flatten, unflatten)


def register_pytree_node(cls, flatten, unflatten):
    _REGISTRY[cls] = (flatten, unflatten)


def register_dataclass(cls):
    names = [f.name for f in dataclasses.fields(cls)]

    def fl(x):
        return [getattr(x, n) for n in names], None

    def unfl(aux, children):
        return cls(**dict(zip(names, children)))
    _REGISTRY[cls] = (fl, unfl)
    return cls


class _Leaf:
    def __repr__(self):
        return "*"


LEAF = _Leaf()


def _flatten(x, leaves, is_leaf):
    if is_leaf is not None and is_leaf(x):
        leaves.append(x)
        return LEAF
    if x is None:
        return ("none",)
    t = type(x)
    if t in _REGISTRY:
        children, aux = _REGISTRY[t][0](x)
        return ("reg", t, aux, [_flatten(c, leaves, is_leaf) for c in children])
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return ("namedtuple", t, [_flatten(c, leaves, is_leaf) for c in x])
    if t is list or t is tuple:
        return (t.__name__, [_flatten(c, leaves, is_leaf) for c in x])
    if isinstance(x, dict):
        keys = sorted(x.keys())
        return ("dict", t, keys, [_flatten(x[k], leaves, is_leaf) for k in keys])
    leaves.append(x)
    return LEAF


def _unflatten(td, it):
    if td is LEAF:
        return next(it)
    tag = td[0]
    if tag == "none":
        return None
    if tag == "reg":
        _, t, aux, ch = td
        return _REGISTRY[t][1](aux, [_unflatten(c, it) for c in ch])
    if tag == "namedtuple":
        return td[1](*[_unflatten(c, it) for c in td[2]])
    if tag == "list":
        return [_unflatten(c, it) for c in td[1]]
    if tag == "tuple":
        return tuple(_unflatten(c, it) for c in td[1])
    if tag == "dict":
        _, t, keys, ch = td
        return {k: _unflatten(c, it) for k, c in zip(keys, ch)}
    raise TypeError(td)


def tree_flatten(tree, is_leaf=None):
    leaves = []
    td = _flatten(tree, leaves, is_leaf)
    return leaves, td


def tree_unflatten(treedef, leaves):
    return _unflatten(treedef, iter(leaves))


def tree_leaves(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[0]


def tree_structure(tree):
    return tree_flatten(tree)[1]


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, td = tree_flatten(tree, is_leaf)
    others = [tree_flatten(r, is_leaf)[0] for r in rest]
    for o in others:
        if len(o) != len(leaves):
            raise ValueError("tree_map: trees do not match")
    return tree_unflatten(td, [f(*xs) for xs in zip(leaves, *others)])


map = tree_map
flatten = tree_flatten
unflatten = tree_unflatten
leaves = tree_leaves
