"""Sub-module import paths of the reference for modules that are single files here.

The reference lays ``mx.module`` / ``mx.symbol`` / ``mx.io`` / ``mx.image`` out as packages (``mxnet/module/base_module.py`` ...), and user code
imports from those paths (``from mxnet.module.base_module import BaseModule``).  In this framework each of them is one module; this helper
registers the reference's sub-module names as real ``sys.modules`` entries that expose the same objects, so both import styles work."""
import sys
import types


def submodule(parent_name, sub, exports=None, getattr_fn=None, doc=None):
    """Create ``<parent>.<sub>`` as a module whose attributes are ``exports`` (a dict) and, lazily, ``getattr_fn(name)``."""
    full = parent_name + "." + sub
    m = types.ModuleType(full, doc)
    if exports:
        m.__dict__.update(exports)
        m.__all__ = [k for k in exports if not k.startswith("_")]
    if getattr_fn is not None:
        m.__getattr__ = getattr_fn
    sys.modules[full] = m
    parent = sys.modules.get(parent_name)
    if parent is not None and not hasattr(parent, sub):
        setattr(parent, sub, m)
    return m
