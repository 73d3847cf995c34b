"""TEST INFRASTRUCTURE ONLY -- recipe that installs the UNMODIFIED reference (intel/rho-diffusion) into ``oracle/_ref``.

``/root/reference`` exists only in the build container; the GPU box gets the working tree (git-ignored files included), so
an installed copy under ``oracle/_ref`` (listed in ``.gitignore``, NOT in ``.gpurunignore``) is what lets ``bench.py --impl
reference`` and the ``cpu_baseline`` leg time the reference's own modules there (``cpu_baseline.kind = "reference"``).
The install is the stock ``pip install --no-deps --target`` of the reference's own ``pyproject.toml`` from a scratch copy
(``/root/reference`` is read-only and setuptools writes ``build/`` next to the sources); no reference source enters the git
history.  Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present.

    python oracle/build_ref.py            # -> oracle/_ref/rho_diffusion (+ dist-info)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SRC = os.environ.get("RHO_REFERENCE_ROOT", "/root/reference")


def build(force: bool = False) -> bool:
    """Returns True when ``oracle/_ref/rho_diffusion`` exists afterwards."""
    pkg = os.path.join(TARGET, "rho_diffusion")
    if os.path.isdir(pkg) and not force:
        return True
    if not os.path.isdir(os.path.join(SRC, "rho_diffusion")):
        return False
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", TARGET, work]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout[-2000:] + res.stderr[-4000:])
            return False
    return os.path.isdir(pkg)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ok" if ok else "unavailable")
    sys.exit(0 if ok else 1)
