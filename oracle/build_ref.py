#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference, made available to the GPU box.

TEST INFRASTRUCTURE.  pySDC is pure Python; "building" it means making its package importable where
``/root/reference`` does not exist.  This recipe copies the reference's own ``pySDC/{core,helpers,implementations}``
(1.7 MB) plus its tutorials and the few reference tests that exercise the classes on the path (0.4 MB; projects and
playgrounds are not needed) byte for byte from
where they lie under ``/root/reference`` into ``oracle/_ref/pySDC``.  ``oracle/_ref/`` is git-ignored (reference sources
never enter the history) but not gpurun-ignored, so the copy travels to the GPU box like a built ``.so``.  A manifest
with one SHA-256 per file is written next to it; ``verify()`` re-checks it, which is how the tests on the GPU box know
the tree is the reference as it was in the build container.

Who may import ``oracle/_ref``: ``tests/`` (the reference's controller driving the CUDA plug-in classes),
``bench.py --impl reference`` / ``cpu_baseline`` (the reference's CPU path timed on the box's host cores), and the e2e
leg of ``bench.py`` for the CONTROLLER only (the reference's ``controller_nonMPI`` around this repo's kernels — exactly
the drop-in a pySDC user would run).  The third-party ``qmat`` package is absent from the image; ``oracle/qmat_shim``
(committed, our own restatement) stands in for it on both sides.

    python oracle/build_ref.py            # copy + manifest (build container)
    python oracle/build_ref.py verify     # anywhere
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PYSDC_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
PARTS = ["__init__.py", "core", "helpers", "implementations",
         # the reference's own tutorials and the tests of them / of the classes on the path: tests/test_reference_suite.py
         # runs them, unmodified, on the plug-in classes
         "tutorial", "tests/__init__.py", "tests/test_tutorials", "tests/test_transfer_classes/test_mesh_to_mesh.py",
         "tests/test_2d_fd_accuracy.py", "tests/test_convergence_controllers/test_check_convergence.py",
         "tests/test_sweepers/test_MPI_sweeper.py", "tests/test_transfer_classes/test_base_transfer_MPI.py",
         # the reference's own CPU-vs-GPU check of its CuPy heat class (projects/GPU/heat.py:61-94)
         "projects/__init__.py", "projects/GPU/__init__.py", "projects/GPU/heat.py"]
MANIFEST = os.path.join(DST, "MANIFEST.json")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _walk(root):
    for base, dirs, files in os.walk(root):
        dirs[:] = sorted(d for d in dirs if d != "__pycache__")
        for name in sorted(files):
            if name.endswith(".pyc"):
                continue
            yield os.path.relpath(os.path.join(base, name), root)


def build():
    """Copy the reference package where it lies; returns the destination or None when the source tree is absent."""
    src = os.path.join(REF_SRC, "pySDC")
    if not os.path.isdir(src):
        return DST if os.path.isfile(MANIFEST) else None
    pkg = os.path.join(DST, "pySDC")
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(pkg)
    for part in PARTS:
        s, d = os.path.join(src, part), os.path.join(pkg, part)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "data"))
        else:
            shutil.copy2(s, d)
    for extra in ("LICENSE",):
        if os.path.isfile(os.path.join(REF_SRC, extra)):
            shutil.copy2(os.path.join(REF_SRC, extra), os.path.join(DST, extra))
    files = {rel: _sha(os.path.join(pkg, rel)) for rel in _walk(pkg)}
    with open(MANIFEST, "w") as f:
        json.dump(dict(source=src, parts=PARTS, files=files), f, indent=0, sort_keys=True)
    return DST


def verify():
    """True when every file of oracle/_ref/pySDC matches the manifest written at copy time (nothing edited)."""
    if not os.path.isfile(MANIFEST):
        return False
    with open(MANIFEST) as f:
        man = json.load(f)
    pkg = os.path.join(DST, "pySDC")
    have = set(_walk(pkg))
    return have == set(man["files"]) and all(_sha(os.path.join(pkg, rel)) == h for rel, h in man["files"].items())


def reference_paths():
    """sys.path entries that make ``import pySDC`` (the unmodified reference) and ``import qmat`` (the stand-in) work:
    the mounted tree in the build container, else the shipped copy; None when neither exists."""
    shim = os.path.join(HERE, "qmat_shim")
    if os.path.isdir(os.path.join(REF_SRC, "pySDC")):
        return [shim, REF_SRC]
    if os.path.isfile(MANIFEST):
        return [shim, DST]
    return None


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "verify":
        ok = verify()
        print("oracle/_ref verified" if ok else "oracle/_ref missing or modified")
        sys.exit(0 if ok else 1)
    print(build())
