"""Importable name of the product package.

The package directory is `fest-3d_b200/` (the name the build contract fixes); a hyphen cannot appear in an import statement,
so this module presents that directory as the package `fest3d_b200`: `import fest3d_b200.solver`, `from fest3d_b200 import capi`.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "fest-3d_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _f.name, "exec"))
