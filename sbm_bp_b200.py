"""Import alias: the package directory is named ``sbm-bp_b200`` (not a valid Python identifier).

``import sbm_bp_b200`` resolves to this module, which turns itself into a package rooted at that
directory, so ``from sbm_bp_b200 import api`` and friends work from the repo root.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "sbm-bp_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
