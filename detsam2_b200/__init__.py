"""Import alias: the package sources live in ``det-sam2_b200/`` (not a valid Python identifier);
this stub makes them importable as ``detsam2_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "det-sam2_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
