"""Small object-plumbing helpers with the reference's names (open_flamingo/src/utils.py:1-31)."""
from functools import reduce


def extend_instance(obj, mixin):
    """Give an existing instance an extra base class, placed FIRST in the MRO so the mixin's forward()
    wraps the original one (what FlamingoLMMixin relies on; reference utils.py:1-7)."""
    cls = type(obj)
    obj.__class__ = type(cls.__name__, (mixin, cls), {})


def getattr_recursive(obj, att):
    """getattr_recursive(m, "a.b.c") -> m.a.b.c ; the empty path returns the object itself."""
    if not att:
        return obj
    return reduce(getattr, att.split("."), obj)


def setattr_recursive(obj, att, val):
    """setattr_recursive(m, "a.b.c", v) sets m.a.b.c = v."""
    head, _, leaf = att.rpartition(".")
    setattr(getattr_recursive(obj, head), leaf, val)


def apply_with_stopping_condition(module, apply_fn, apply_condition=None, stopping_condition=None, **other_args):
    """Depth-first walk applying `apply_fn` where `apply_condition` holds, pruning subtrees at
    `stopping_condition` (reference utils.py:34-48; used there only by the FSDP wrapper)."""
    if stopping_condition is not None and stopping_condition(module):
        return
    if apply_condition is None or apply_condition(module):
        apply_fn(module, **other_args)
    for child in module.children():
        apply_with_stopping_condition(child, apply_fn, apply_condition=apply_condition,
                                      stopping_condition=stopping_condition, **other_args)
