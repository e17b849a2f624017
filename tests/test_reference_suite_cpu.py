"""Runs the reference's OWN test scripts, unchanged, against this repository's `models` package (CPU).

Only possible where the reference tree is mounted (/root/reference, the build container); skipped elsewhere.
tests/test_numerical_validation.py imports matplotlib at module level, which is not installed, so a two-file
stub is put on PYTHONPATH for it (the tests never plot unless asked).
"""
import os
import subprocess
import sys

import pytest

from conftest import REPO

REF = os.environ.get("UMNN_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="reference tree not mounted")


def _run(script, first_path, extra_path, cwd):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([first_path] + extra_path)
    env["CUDA_VISIBLE_DEVICES"] = ""
    return subprocess.run([sys.executable, os.path.join(REF, "tests", script)], capture_output=True, text=True,
                          env=env, cwd=cwd, timeout=900)


def _marked(out):
    """The verdict lines of the reference's print-style tests."""
    return [ln.strip() for ln in out.splitlines() if ("\u2713" in ln or "\u2717" in ln)]


def test_reference_test_jit_passes_against_our_package(tmp_path):
    r = _run("test_jit.py", REPO, [], str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "All tests completed successfully" in r.stdout
    assert "\u2717" not in r.stdout


def test_reference_numerical_validation_gives_the_same_verdicts_as_the_reference(tmp_path):
    stub = tmp_path / "stubs" / "matplotlib"
    stub.mkdir(parents=True)
    (stub / "__init__.py").write_text("def use(*a, **k):\n    pass\n")
    (stub / "pyplot.py").write_text(
        "def __getattr__(name):\n    def _noop(*a, **k):\n        return None\n    return _noop\n")
    ours = _run("test_numerical_validation.py", REPO, [str(tmp_path / "stubs")], str(tmp_path))
    theirs = _run("test_numerical_validation.py", REF, [str(tmp_path / "stubs")], str(tmp_path))
    assert ours.returncode == 0, ours.stdout[-2000:] + ours.stderr[-2000:]
    assert theirs.returncode == 0
    # same checks, same verdicts, same printed numbers (known answers 14/3, 6, 2, 26/3, e-1; the y = x^3 fit)
    assert _marked(ours.stdout) == _marked(theirs.stdout)
    assert any("Successfully fitted" in ln for ln in _marked(ours.stdout))
