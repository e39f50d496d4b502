#!/usr/bin/env python
"""Stage the UNMODIFIED reference under oracle/_ref/ so it travels to the GPU box.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference (Meatssauce/alpha-zero-general-inflexion) is pure Python: there is nothing to compile, so the
"build" of oracle/_ref is a verbatim copy of the files the self-play path imports, made from where they lie under
/root/reference (read-only, present in the build container only).  oracle/_ref/ is git-ignored (no reference source
enters the history) but NOT gpurun-ignored, so ``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the
reference's own ``MCTS`` + ``InflexionGame`` + ``NNetWrapper`` on the GPU box's host cores (BASELINE.md section 3,
SURVEY.md 8d).  A manifest with the SHA-256 of every file is written next to the copies; bench.py re-hashes the files
before it runs them and refuses a copy that was edited.

    python oracle/make_ref.py [--src /root/reference]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
# what MCTS.search / getActionProb / Coach.executeEpisode / Arena.playGame import (MCTS.py:1-9, Coach.py:1-19)
FILES = ["MCTS.py", "Game.py", "flags.py", "utils.py", "NeuralNet.py", "Coach.py", "Arena.py", "LICENSE",
         "inflexion/__init__.py", "inflexion/InflexionGame.py", "inflexion/InflexionPlayers.py", "inflexion/utils.py",
         "inflexion/pytorch/__init__.py", "inflexion/pytorch/InflexionNNet.py", "inflexion/pytorch/NNet.py"]


def sha256(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(src: str = "/root/reference", dest: str = DEST) -> dict:
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = sha256(d)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def verify(dest: str = DEST) -> bool:
    """True when every staged file still has the hash recorded at staging time (i.e. the copy is unmodified)."""
    try:
        files = json.load(open(os.path.join(dest, "MANIFEST.json")))["files"]
        return set(files) == set(FILES) and all(sha256(os.path.join(dest, rel)) == h for rel, h in files.items())
    except Exception:
        return False


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    m = stage(a.src)
    print(f"[make_ref] {len(m)} files -> {DEST} (verified: {verify()})")
