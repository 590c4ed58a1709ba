"""Recipe for `oracle/_ref/`: the UNMODIFIED hot-path modules of the reference, staged so that they can travel
to the GPU box -- TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/make_ref.py [--reference /root/reference]

The reference is pure Python (no build system, no native code).  Its hot path is four files
(`src/gpr.py`, `src/tools/uncertainty_prop.py`, `src/dynamics.py`, `src/mpc.py`); this script
copies them byte for byte from the read-only reference checkout into `oracle/_ref/src/...` and writes a
three-line stand-in for the one import that cannot be satisfied offline (`import cyipopt` at `src/mpc.py:4`;
only `get_optimal_trajectory` touches the module, `src/mpc.py:298`).  Nothing is edited and nothing is
committed: `oracle/_ref/` is git-ignored (it is NOT gpurun-ignored, so `bench.py --impl reference` and the
`cpu_baseline` leg can time the reference's own code on the GPU box's host cores, where `/root/reference` does
not exist).  `__graft_entry__.build()` runs this whenever the reference checkout is present.

`MANIFEST.json` records the sha256 of every staged file so a reader can check that the timed code is the
reference's.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
# the hot path + the package markers that make `import src....` resolve
FILES = [
    "src/__init__.py",
    "src/gpr.py",
    "src/dynamics.py",
    "src/mpc.py",
    "src/tools/__init__.py",
    "src/tools/uncertainty_prop.py",
]
CYIPOPT_STUB = '''"""Stand-in for the cyipopt package (not installable offline).  The reference imports it at module top
(`src/mpc.py:4`) but only `get_optimal_trajectory` uses it; objective / gradient / cost never do."""
STUB = True
'''


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(reference="/root/reference", dest=DEST, quiet=False):
    """Copy the hot-path modules; returns the manifest dict, or None if the reference checkout is absent."""
    if not os.path.isdir(os.path.join(reference, "src")):
        return None
    manifest = {"reference": reference, "files": {}}
    for rel in FILES:
        src = os.path.join(reference, rel)
        dst = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest["files"][rel] = sha256(dst)
    stub_dir = os.path.join(dest, "_stubs", "cyipopt")
    os.makedirs(stub_dir, exist_ok=True)
    with open(os.path.join(stub_dir, "__init__.py"), "w") as f:
        f.write(CYIPOPT_STUB)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"staged {len(FILES)} reference files into {dest}")
    return manifest


def available(dest=DEST):
    return all(os.path.exists(os.path.join(dest, rel)) for rel in FILES)


def import_reference(dest=DEST):
    """Put the staged tree on sys.path (and the cyipopt stand-in only if the real package is missing) and return
    the reference's modules: (gpr, uncertainty_prop, dynamics, mpc)."""
    if not available(dest):
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists")
    if dest not in sys.path:
        sys.path.insert(0, dest)
    try:
        import cyipopt  # noqa: F401
    except ImportError:
        sys.path.insert(0, os.path.join(dest, "_stubs"))
    import importlib
    mods = [importlib.import_module(m) for m in ("src.gpr", "src.tools.uncertainty_prop", "src.dynamics", "src.mpc")]
    for m in mods:
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(dest)), m.__file__
    return tuple(mods)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    if stage(a.reference) is None:
        sys.exit(f"no reference checkout at {a.reference}")
