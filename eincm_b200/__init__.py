"""Importable alias for the package directory ``edge-informed-contrast-maximization_b200/``.

The package directory carries the repository's name (with hyphens), which Python cannot import directly;
this shim puts that directory on the package search path so that ``import eincm_b200.losses`` etc. resolve
to the modules that live there.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         'edge-informed-contrast-maximization_b200')
__path__.insert(0, _PKG_DIR)
with open(_os.path.join(_PKG_DIR, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, '__init__.py'), 'exec'))
del _f
