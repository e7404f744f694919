"""Import alias for the ``open-kinematics_b200/`` package directory.

The repository layout names the package directory ``open-kinematics_b200``;
a hyphen is not importable, so this module turns itself into that package
(``__path__`` points at the directory and its ``__init__`` is executed here).
``import open_kinematics_b200.core.solver`` therefore resolves to
``open-kinematics_b200/core/solver.py``.
"""

import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "open-kinematics_b200")
__path__ = [_PKG_DIR]
__file__ = _os.path.join(_PKG_DIR, "__init__.py")
with open(__file__, "r", encoding="utf-8") as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
