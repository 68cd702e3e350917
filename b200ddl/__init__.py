"""Import alias: ``import b200ddl`` loads the package that lives in ``distributed-deep-learning-workshop_b200/``.

The on-disk directory keeps the name the build brief asks for (it contains a hyphen, so it cannot be imported
directly); this shim points the ``b200ddl`` package at that directory and runs its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "distributed-deep-learning-workshop_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
