"""Small object-plumbing helpers with the reference's names (open_flamingo/src/utils.py:1-31).

`apply_with_stopping_condition` (utils.py:34-48) is not provided: its only caller is the FSDP wrapper
(flamingo.py:202-301), which is out of scope here."""
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
