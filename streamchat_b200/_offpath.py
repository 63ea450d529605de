"""Names the reference driver imports next to the hot-path ones but that are OUTSIDE the path this package
serves (optical-flow / SSIM frame filters, PCA plots, the LLaMA-3 client, non-multimodal tree
search ...).  The drop-in modules must still export them — `from utiles import compute_gradients, ...` at
inference_streaming_longva_v2.py:19-26 is one statement — so each shadow module resolves such a name lazily:

  1. from the reference's own module of the same name, when the reference tree is further down `sys.path`
     (the documented drop-in run puts dropin/ first and the reference tree last) and its dependencies import;
  2. otherwise as a placeholder that raises NotImplementedError when CALLED (never at import), naming the
     reference file to take it from.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from typing import Dict, Optional

_ref_modules: Dict[str, object] = {}


def _reference_module(dotted: str, shadow_file: str):
    """The module `dotted` as the reference tree defines it (the first sys.path entry, other than the drop-in
    directory that holds `shadow_file`, that contains it), loaded under a private name.  None when absent or when
    its own imports fail (e.g. matplotlib / gradio missing)."""
    if dotted in _ref_modules:
        return _ref_modules[dotted]
    mod = None
    rel = dotted.replace(".", os.sep)
    shadow_root = os.path.abspath(shadow_file)
    for _ in dotted.split("."):
        shadow_root = os.path.dirname(shadow_root)
    for entry in list(sys.path):
        root = os.path.abspath(entry or ".")
        if root == shadow_root:
            continue
        for cand in (os.path.join(root, rel + ".py"), os.path.join(root, rel, "__init__.py")):
            if os.path.exists(cand):
                try:
                    spec = importlib.util.spec_from_file_location("_streamchat_reference_." + dotted, cand)
                    m = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(m)
                    mod = m
                except Exception:       # the reference module's own dependencies are missing: placeholders serve
                    mod = None
                break
        if mod is not None:
            break
    _ref_modules[dotted] = mod
    return mod


def placeholder(name: str, where: str):
    def _off_path(*args, **kwargs):
        raise NotImplementedError(
            f"{name} ({where}) is outside the hot path streamchat_b200 serves (SURVEY.md §8); put the reference "
            f"tree on PYTHONPATH after dropin/ to use the reference's own implementation")
    _off_path.__name__ = _off_path.__qualname__ = name
    _off_path.__doc__ = f"Placeholder for the reference's {name} ({where}); raises when called."
    return _off_path


def resolver(dotted: str, shadow_file: str, off_path_names: Dict[str, str]):
    """Module-level `__getattr__` (PEP 562) for a shadow module: `off_path_names` maps name -> reference file:line."""
    def __getattr__(name: str):
        if name.startswith("__"):
            raise AttributeError(name)
        ref = _reference_module(dotted, shadow_file)
        if ref is not None and hasattr(ref, name):
            return getattr(ref, name)
        if name in off_path_names:
            return placeholder(name, off_path_names[name])
        raise AttributeError(f"module {dotted!r} (streamchat_b200 drop-in) has no attribute {name!r}")
    return __getattr__
