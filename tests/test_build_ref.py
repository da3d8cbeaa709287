"""oracle/build_ref.py stages byte-identical copies of the reference's modules (build container only)."""
import hashlib
import json
import os

import pytest

from oracle import build_ref


def test_staged_reference_files_are_byte_identical():
    if not build_ref.available():
        pytest.skip("/root/reference not present (GPU box)")
    assert build_ref.build()
    p = build_ref.staged_path()
    assert p is not None
    man = json.load(open(os.path.join(p, "MANIFEST.json")))
    assert "models/ddpm/unet.py" in man["files"] and "losses/ddpm.py" in man["files"]
    for rel, sha in man["files"].items():
        assert hashlib.sha256(open(os.path.join(p, rel), "rb").read()).hexdigest() == sha
        assert hashlib.sha256(open(os.path.join(build_ref.REF, rel), "rb").read()).hexdigest() == sha


def test_staged_reference_is_not_tracked_by_git():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gi = open(os.path.join(root, ".gitignore")).read()
    assert "oracle/_ref/" in gi
