"""Loader for the UNMODIFIED reference extension built into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  ``load()`` returns the reference's own ``UMPA.model``
extension module (classes UMPAModelNoDF / UMPAModelDF / UMPAModelDFKernel and the
hooks spm / spmq / test_CostArgsDFKernel ...), or None when it was never built
(``oracle/build.py`` needs /root/reference, which only exists in the build
container; the prebuilt .so travels to the GPU box).

``import UMPA`` itself is not possible here (UMPA/__init__.py pulls align.py ->
matplotlib), so the extension is loaded directly from its file.
"""
import importlib.machinery
import importlib.util
import sys

from . import build as _build

_mod = None


def load(build_if_missing=True):
    global _mod
    if _mod is not None:
        return _mod
    path = _build.REF_SO
    if not _build.ref_available():
        if not build_if_missing:
            return None
        try:
            path = _build.build_ref()
        except Exception:
            return None
        if path is None:
            return None
    name = "UMPA.model"
    loader = importlib.machinery.ExtensionFileLoader(name, path)
    spec = importlib.util.spec_from_file_location(name, path, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    try:
        loader.exec_module(mod)
    except Exception:
        return None
    sys.modules.setdefault(name, mod)
    _mod = mod
    return _mod
