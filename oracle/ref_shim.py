"""Import the UNMODIFIED reference (read-only at /root/reference) in the build container.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (never on the GPU box): it is used by
``make_golden.py`` to generate the golden vectors under tests/golden/ and by the ``needs_reference`` tests that
re-check the oracle against the live reference.  The reference's scripts import packages that are absent here
(timm, matplotlib, pymilvus, open_clip ...); none of them is touched by the hot-path functions, so they are
replaced by empty stub modules.  No reference file is modified or copied.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("KNN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "test.py"))


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _stub(name: str, attrs=()):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []  # behave like a package so `import a.b` works
    for a in attrs:
        setattr(m, a, _Anything)

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything

    m.__getattr__ = _getattr  # type: ignore[attr-defined]
    sys.modules[name] = m
    return m


_ready = False


def install() -> None:
    global _ready
    if _ready:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    import torch  # noqa: F401
    import torchvision  # noqa: F401
    try:
        import transformers  # noqa: F401  (probes find_spec("timm") before the stub exists)
    except Exception:
        pass
    for name in ("timm", "timm.data", "timm.models", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
                 "pymilvus", "open_clip", "faiss", "boto3", "onnxruntime", "wfdb", "dotenv", "seaborn"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _ready = True


def module(name: str):
    """e.g. module('test'), module('train'), module('evaluate_nih_zilliz'), module('fusion_eval.metrics')."""
    install()
    saved = sys.modules.get(name)
    if saved is not None and not getattr(saved, "__file__", "").startswith(REFERENCE_ROOT):
        del sys.modules[name]  # e.g. the stdlib `test` package shadows the reference's test.py
    return importlib.import_module(name)
