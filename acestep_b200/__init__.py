"""Importable alias for the package directory `ace-step-1.5-for-windows_b200/` (whose name is not
a valid Python identifier).  `import acestep_b200.dit` resolves to files in that directory."""
import os as _os

_PKG = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "ace-step-1.5-for-windows_b200")
__path__ = [_PKG]
with open(_os.path.join(_PKG, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG, "__init__.py"), "exec"))
