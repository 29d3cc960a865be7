#!/usr/bin/env python
"""Install the UNMODIFIED reference into the git-ignored baseline/_ref/ for bench.py's `--impl reference` arm.

    python scripts/install_reference.py            # build container only (/root/reference exists here)

Recipe (the task contract's one allowed offline install):
  1. `python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref <src>`
     from a /tmp copy of /root/reference (the source tree is read-only).  The reference ships no setup.py /
     pyproject.toml (it is a set of scripts), so pip refuses it -- the outcome is recorded.
  2. Fallback = what such an install would have placed in the target: the reference's importable packages
     (`jen1/`, `utils/`), `.py` files only, byte-identical (sha256 recorded in baseline/_ref/INSTALL.json).

baseline/_ref is listed in .gitignore (never enters history: reference sources are not copied into the repo) but NOT
in .gpurunignore, so it travels to the GPU box with the snapshot like the built .so files.  Nothing in the product path
imports it; only `bench.py --impl reference` does (through oracle/ref_import.py's two import shims).
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("jen1", "utils")


def main() -> int:
    if not os.path.isdir(SRC):
        print("install_reference: %s not present (GPU box?) -- nothing to do" % SRC)
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    record = {"source": SRC, "pip": None, "method": None, "files": {}}
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", DST, work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        record["pip"] = {"rc": r.returncode, "tail": (r.stderr or r.stdout).strip().splitlines()[-3:]}
        if r.returncode == 0 and os.path.isdir(os.path.join(DST, "jen1")):
            record["method"] = "pip install --target"
        else:
            record["method"] = "package copy (pip refused: no setup.py / pyproject.toml in the reference)"
            for pkg in PACKAGES:
                shutil.copytree(os.path.join(work, pkg), os.path.join(DST, pkg))
    for base, _, files in os.walk(DST):
        for fn in files:
            if fn.endswith(".py"):
                p = os.path.join(base, fn)
                rel = os.path.relpath(p, DST)
                h = hashlib.sha256(open(p, "rb").read()).hexdigest()
                src = os.path.join(SRC, rel)
                record["files"][rel] = {"sha256": h, "identical_to_source": os.path.exists(src) and
                                        hashlib.sha256(open(src, "rb").read()).hexdigest() == h}
    assert all(v["identical_to_source"] for v in record["files"].values()), "baseline/_ref differs from the reference"
    with open(os.path.join(DST, "INSTALL.json"), "w") as f:
        json.dump(record, f, indent=1)
    print("install_reference: %s -> %s (%d files, %s)" % (SRC, DST, len(record["files"]), record["method"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
