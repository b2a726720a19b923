"""TEST INFRASTRUCTURE ONLY -- loader for the upstream reference scripts.

Executes only the ``import`` / ``def`` / ``class`` top-level statements of a reference
script (e.g. /root/reference/CRVAE_lorenz96.py) so that its model classes and trainers can be
called without running the script's module-level training driver
(CRVAE_lorenz96.py:730-796) and without its unavailable imports (tensorflow :8, matplotlib :6).

The reference tree only exists in the build container; nothing that runs on the GPU box
(pytest -m gpu, smoke(), bench.py) may call this.  It is used by tests/golden/make_golden.py to
produce the committed golden vectors and by CPU tests that are skipped when the tree is absent.
"""
from __future__ import annotations

import ast
import os
import types

REFERENCE_ROOT = os.environ.get("CRVAE_REFERENCE_ROOT", "/root/reference")
_DROP_IMPORT_ROOTS = ("tensorflow", "matplotlib")


class _NoOpPlt:
    """Stand-in for matplotlib.pyplot: every attribute is a callable returning another no-op."""

    def __getattr__(self, name):
        return _NoOpPlt()

    def __call__(self, *a, **k):
        return _NoOpPlt()

    def __iter__(self):
        return iter((_NoOpPlt(), _NoOpPlt()))


def reference_available(script: str = "CRVAE_lorenz96.py") -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, script))


def load_reference(script: str = "CRVAE_lorenz96.py") -> types.SimpleNamespace:
    """Return a namespace holding the definitions of a reference script."""
    path = os.path.join(REFERENCE_ROOT, script)
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    future, keep = [], []
    for node in tree.body:
        if isinstance(node, ast.ImportFrom) and node.module == "__future__":
            future.append(node)
        elif isinstance(node, (ast.Import, ast.ImportFrom)):
            roots = ([a.name.split(".")[0] for a in node.names] if isinstance(node, ast.Import)
                     else [(node.module or "").split(".")[0]])
            if any(r in _DROP_IMPORT_ROOTS for r in roots):
                continue
            keep.append(node)
        elif isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            keep.append(node)
    mod = ast.Module(body=future + keep, type_ignores=[])
    ns: dict = {"__name__": "reference_" + script.replace("-", "_").replace(".py", ""),
                "plt": _NoOpPlt()}
    exec(compile(mod, path, "exec"), ns)
    # sklearn>=1.9 rejects TSNE(n_iter=...) (CRVAE_lorenz96.py:435); plotting is out of scope.
    if "visualization" in ns:
        ns["visualization"] = lambda *a, **k: None
    return types.SimpleNamespace(**ns)
