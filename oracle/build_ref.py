"""Recipe: compile the reference's own hot-path sources into ``oracle/_ref``.

Reads (never writes) ``/root/reference``; every output (generated C and the
extension modules) goes to ``oracle/_ref/``.  No reference source is copied
into the repository: Cython is pointed at the files where they lie and only
the compiled artefacts are kept.  ``oracle/_ref`` is git-ignored and is NOT
listed in ``.gpurunignore``, so the ``.so`` files travel to the GPU box where
``/root/reference`` does not exist.

    python oracle/build_ref.py [--force]

Sources compiled (unmodified):
    torchreid/metrics/distance.py            -> _ref/distance.*.so
    torchreid/metrics/rank.py                -> _ref/rank.*.so
    torchreid/utils/rerank.py                -> _ref/rerank.*.so
    torchreid/metrics/rank_cylib/rank_cy.pyx -> _ref/rank_cy.*.so
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("IEEE_REFERENCE_ROOT", "/root/reference")

SOURCES = {
    "distance": "torchreid/metrics/distance.py",
    "rank": "torchreid/metrics/rank.py",
    "rerank": "torchreid/utils/rerank.py",
    "rank_cy": "torchreid/metrics/rank_cylib/rank_cy.pyx",
}


def ext_path(name: str) -> str:
    return os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))


def reference_present() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_ROOT, p)) for p in SOURCES.values())


def built() -> bool:
    return all(os.path.isfile(ext_path(n)) for n in SOURCES)


def build(force: bool = False, verbose: bool = True) -> bool:
    """Build every module; returns True when all four exist afterwards."""
    if not reference_present():
        if verbose:
            print(f"[oracle/_ref] {REFERENCE_ROOT} not present; using prebuilt files only")
        return built()
    import numpy as np

    os.makedirs(OUT, exist_ok=True)
    py_inc = sysconfig.get_paths()["include"]
    for name, rel in SOURCES.items():
        src = os.path.join(REFERENCE_ROOT, rel)
        so = ext_path(name)
        if not force and os.path.isfile(so) and os.path.getmtime(so) >= os.path.getmtime(src):
            continue
        c_file = os.path.join(OUT, name + ".c")
        subprocess.run([sys.executable, "-m", "cython", "-3", "-o", c_file, src], check=True,
                       stdout=subprocess.DEVNULL if not verbose else None)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-w", f"-I{py_inc}", f"-I{np.get_include()}",
                        c_file, "-o", so], check=True)
        os.remove(c_file)  # keep only the compiled artefact
        if verbose:
            print(f"[oracle/_ref] built {os.path.relpath(so, HERE)} from {src}")
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
