#!/usr/bin/env python
"""Offline install of the UNMODIFIED reference (andreeaiana/newsreclib) into ``baseline/_ref`` for the reference arm
of ``bench.py`` (``--impl reference``).  Run in the build container, where ``/root/reference`` exists; the result is
git-ignored (reference sources never enter this repo's history) but travels to the GPU box with the snapshot.

    python baseline/install_ref.py

1. ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`` -- from a
   copy under /tmp because the build writes ``*.egg-info`` into the source tree and ``/root/reference`` is read-only;
   ``--no-deps`` because the reference pins ``lightning`` / ``torch_geometric`` / ``torchmetrics`` / ``hydra`` ...,
   none of which is in the offline wheelhouse.
2. The reference's ``setup.py`` uses ``find_packages()``, and several of its sub-packages carry a MISNAMED
   ``__init.py__`` (``models/components/encoders/**``, ``models/components/layers``, ``metrics``): setuptools does not
   see them, so the wheel lacks exactly the files of the hot path (``text.py``, ``news.py``, ``attention.py``,
   ``user/nrms.py`` ...).  In the source tree they import fine as namespace packages; to give the installed copy the
   same content, every ``*.py`` of ``/root/reference/newsreclib`` that the wheel left out is copied over byte for byte.
3. The wheel's unrelated top-level ``tests`` / ``configs`` packages are dropped.

Prints what was installed and a SHA-256 over the installed hot-path files.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
HOT = ["models/components/encoders/news/text.py", "models/components/encoders/news/news.py",
       "models/components/encoders/user/nrms.py", "models/components/layers/attention.py",
       "models/components/layers/click_predictor.py", "models/general_rec/nrms_module.py",
       "models/abstract_recommender.py"]


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "newsreclib")):
        print(f"{SRC} is absent: this recipe runs in the build container only", file=sys.stderr)
        return 1
    shutil.rmtree(DST, ignore_errors=True)
    with tempfile.TemporaryDirectory() as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(SRC, copy, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DST, copy]
        print("[install_ref]", " ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    added = 0
    for dirpath, _, files in os.walk(os.path.join(SRC, "newsreclib")):
        for f in files:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(dirpath, f), SRC)
            out = os.path.join(DST, rel)
            if not os.path.exists(out):
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(os.path.join(SRC, rel), out)
                added += 1
    for junk in ("tests", "configs"):
        shutil.rmtree(os.path.join(DST, junk), ignore_errors=True)
    h = hashlib.sha256()
    for rel in HOT:
        a = open(os.path.join(SRC, "newsreclib", rel), "rb").read()
        b = open(os.path.join(DST, "newsreclib", rel), "rb").read()
        assert a == b, f"{rel} differs from the reference source"
        h.update(b)
    n = sum(len(fs) for _, _, fs in os.walk(os.path.join(DST, "newsreclib")))
    print(f"[install_ref] {DST}: {n} files ({added} completed from the source tree: namespace sub-packages the wheel "
          f"skipped); hot-path files identical to {SRC}, sha256 {h.hexdigest()[:16]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
