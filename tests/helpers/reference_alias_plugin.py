"""pytest plugin used by tests/test_reference_suite_cpu.py: lets the REFERENCE'S OWN test files (run from
/root/reference/tests, unmodified) import ``phasespace`` and get phasespace_b200, and gives their helper module the two
third-party names it imports at module level (``tensorflow``, ``tensorflow_probability``) as empty stand-ins -- the
object-model tests never call into them."""
import os
import sys
import types

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import phasespace_b200  # noqa: E402

sys.modules.setdefault("phasespace", phasespace_b200)
for name in ("tensorflow", "tensorflow_probability"):
    if name not in sys.modules:
        try:
            __import__(name)
        except ImportError:
            sys.modules[name] = types.ModuleType(name)
