"""Importable alias of the ``dis-yolo_b200`` package directory (a hyphen is not a valid Python
identifier).  All code lives in ``../dis-yolo_b200``; this module only redirects ``__path__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'dis-yolo_b200')
__path__ = [_real]
__file__ = _os.path.join(_real, '__init__.py')
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, 'exec'))
