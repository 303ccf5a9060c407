"""The reference's OWN object-model tests, run unmodified from /root/reference/tests against phasespace_b200 (aliased as
``phasespace``): all of test_nbody_decay.py and everything in test_chain.py that does not generate events -- naming,
name clashes, children validation, grandchildren detection, the set_children latch, the errors for a childless or
resonant top particle, ``__repr__``.  They need neither a GPU nor TensorFlow.  Skipped where /root/reference does not
exist (the GPU box)."""

import os
import subprocess
import sys

import pytest

REFERENCE_TESTS = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))

# (file, -k expression): everything in these files that does not generate events
SELECTION = [
    ("test_nbody_decay.py", ""),
    ("test_chain.py", "not kstargamma and not k1gamma"),  # those two generate events: ported in test_generate_gpu.py
]


@pytest.mark.skipif(not os.path.isdir(REFERENCE_TESTS), reason="the reference tree is not available here")
@pytest.mark.parametrize("filename,expr", SELECTION)
def test_reference_object_model_tests_pass_unmodified(filename, expr, tmp_path):
    env = dict(os.environ, PYTHONPATH=os.path.join(HERE, "helpers") + os.pathsep + os.environ.get("PYTHONPATH", ""),
               PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-p", "reference_alias_plugin",
           "--basetemp", str(tmp_path / "bt"), os.path.join(REFERENCE_TESTS, filename)]
    if expr:
        cmd += ["-k", expr]
    out = subprocess.run(cmd, cwd="/root/reference", env=env, capture_output=True, text=True, timeout=600)
    tail = (out.stdout + out.stderr)[-3000:]
    assert out.returncode == 0, tail
    assert " passed" in out.stdout and "failed" not in out.stdout, tail
