"""Recipe that vendors the UNMODIFIED Python reference into oracle/_ref/ (git-ignored; it travels to the GPU box with the
snapshot), so that `bench.py --impl reference` and the `cpu_baseline` leg time the reference itself (kind "reference")
instead of the CPU port.  Run in the build container, where /root/reference exists:

    python oracle/make_ref.py            # copies /root/reference/hyperseg/**/*.py -> oracle/_ref/hyperseg/

Nothing under oracle/_ref is committed, imported by hyperseg_b200, or edited.  `load_reference()` is the only entry point
(tests, bench.py's CPU legs): it stubs the one third-party import the model files do not need (`ffmpeg`, imported at
module level by hyperseg/utils/utils.py:9) and returns the reference's model factory for a configuration.
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
REF_SRC = os.environ.get("HYPERSEG_REFERENCE", "/root/reference")


def vendor(src: str = REF_SRC, dst: str = REF_DST) -> bool:
    pkg = os.path.join(src, "hyperseg")
    if not os.path.isdir(pkg):
        return False
    out = os.path.join(dst, "hyperseg")
    if os.path.isdir(out):
        shutil.rmtree(out)
    n = 0
    for root, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(root, f), pkg)
                os.makedirs(os.path.dirname(os.path.join(out, rel)), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(out, rel))
                n += 1
    with open(os.path.join(dst, "SOURCE.txt"), "w") as fh:
        fh.write(f"{n} unmodified .py files of {pkg} (copied by oracle/make_ref.py; not part of the repository)\n")
    return n > 0


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DST, "hyperseg", "models", "hyperseg_v1_0.py"))


def load_reference(module: str):
    """Import hyperseg.models.<module> of the vendored reference (e.g. 'hyperseg_v1_0') and return the module."""
    if not available():
        raise RuntimeError("oracle/_ref is empty: run `python oracle/make_ref.py` where /root/reference exists")
    sys.modules.setdefault("ffmpeg", types.ModuleType("ffmpeg"))
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    return importlib.import_module(f"hyperseg.models.{module}")


if __name__ == "__main__":
    ok = vendor()
    print("vendored the reference into", REF_DST if ok else "(nothing: reference not found)")
    sys.exit(0 if ok else 1)
