"""Importable alias of the package directory `bird-sound-event-detecion_b200/` (a hyphenated name is
not a Python identifier): `import bsed_b200` executes that directory's __init__ with
__path__ pointing at it, so `bsed_b200.models.CRNN` etc. resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "bird-sound-event-detecion_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
