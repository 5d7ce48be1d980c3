#!/usr/bin/env python
"""Stage the reference's OWN test-suite, unmodified, so that it can run against the import shim on a B200.

    python tests/ref_suite/fetch.py [/root/reference]

packs ``<reference>/tests/*.py`` verbatim into ``tests/ref_suite/_ref.tar`` (git-ignored: reference sources never
enter this repository's history; the archive travels to the GPU box with the working tree like the built ``.so``).
``test_convergence.py`` is left out: it only draws a matplotlib figure.
``tests/test_gpu_ref_suite.py`` unpacks the archive into a temporary directory and runs pytest on it with
``PYTHONPATH`` = this repository, so that ``from nbmf_mm import NBMF`` resolves to ``nbmf_mm/`` (the shim over
``nbmf_mm_b200``) in float64 parity mode -- the estimator's default dtype.  ``__graft_entry__.build()`` runs this script
when the reference tree is present."""
import sys
import tarfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ARCHIVE = HERE / "_ref.tar"
SKIP = {"test_convergence.py"}


def fetch(reference="/root/reference"):
    src = Path(reference) / "tests"
    if not src.is_dir():
        return []
    names = []
    with tarfile.open(ARCHIVE, "w") as tar:
        for f in sorted(src.glob("*.py")):
            if f.name in SKIP:
                continue
            tar.add(f, arcname=f.name)
            names.append(f.name)
        ini = Path(reference) / "pytest.ini"
        if ini.is_file():
            tar.add(ini, arcname="pytest.ini")
    return names


if __name__ == "__main__":
    names = fetch(*(sys.argv[1:2]))
    print(f"staged {len(names)} reference test files in {ARCHIVE}: {' '.join(names)}")
