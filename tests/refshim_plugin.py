"""pytest plugin (TEST INFRASTRUCTURE): `-p refshim_plugin` installs the GPy shim of tests/refshim.py and routes the
product's engine to the CPU oracle before collection, so that the reference's own unit tests
(/root/reference/src/tests/unit/autoks/*.py, unmodified) can be collected and run against ms_project_b200.gpy_compat."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refshim  # noqa: E402

refshim.install()
import oracle_engine  # noqa: E402

_ctx = oracle_engine.installed()
_ctx.__enter__()
