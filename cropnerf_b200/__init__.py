"""Importable alias of the product package.

The package directory is named after the reference repository
(``cropnerf-a-neural-radiance-field-based-framework_b200/``), which is not a valid Python identifier; this stub makes
``import cropnerf_b200`` (and ``cropnerf_b200.<module>``) resolve to the modules in that directory.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "cropnerf-a-neural-radiance-field-based-framework_b200")
__path__ = [_pkg_dir]
with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
