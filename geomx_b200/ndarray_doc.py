"""Docstring assembly for operator functions (reference: ``python/mxnet/ndarray_doc.py`` / ``symbol_doc.py``: the text the generated operator
wrappers carry).  Operators are hand-written Python functions here, so this is only used by ``tools/gen_api_doc.py``-style listings and by
code that builds wrappers programmatically."""
import re

__all__ = ["NDArrayDoc", "SymbolDoc", "_build_doc"]


class NDArrayDoc:
    """Base class for extra documentation attached to an operator: sub-class it as ``<op name>Doc`` and put examples in the docstring."""


class SymbolDoc:
    """Same for symbolic operators; ``get_output_shape(sym, **input_shapes)`` is the helper the reference's examples use."""

    @staticmethod
    def get_output_shape(sym, **input_shapes):
        _, out, _ = sym.infer_shape(**input_shapes)
        return dict(zip(sym.list_outputs(), out))


def _build_doc(func_name, desc, arg_names, arg_types, arg_desc, key_var_num_args=None, ret_type=None):
    """``desc`` + a numpy-style Parameters section built from the three parallel lists (+ Returns)."""
    lines = []
    for name, typ, d in zip(arg_names, arg_types, arg_desc):
        if key_var_num_args and name == key_var_num_args:
            continue
        lines.append("%s : %s" % (name, typ))
        if d:
            lines.append("    " + re.sub(r"\s+", " ", d).strip())
    out = "%s\n\nParameters\n----------\n%s\n" % (desc.strip(), "\n".join(lines))
    if ret_type:
        out += "\nReturns\n-------\nout : %s\n    The result of ``%s``.\n" % (ret_type, func_name)
    return out
