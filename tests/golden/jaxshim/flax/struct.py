"""flax.struct stand-in: frozen dataclass + .replace, registered as a pytree node (field order = flatten order)."""
import dataclasses

from jax import tree_util


def field(pytree_node=True, **kw):
    md = dict(kw.pop("metadata", {}) or {})
    md["pytree_node"] = pytree_node
    return dataclasses.field(metadata=md, **kw)


def dataclass(cls=None, **kw):
    if cls is None:
        return lambda c: dataclass(c, **kw)
    dc = dataclasses.dataclass(frozen=True)(cls)

    def replace(self, **updates):
        return dataclasses.replace(self, **updates)
    dc.replace = replace
    tree_util.register_dataclass(dc)
    return dc


class PyTreeNode:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        dataclass(cls)
