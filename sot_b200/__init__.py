"""Import alias for the package directory `1d-spectral-optimal-transport_b200/` (not a valid
Python identifier).  `import sot_b200` exposes that directory's modules: `sot_b200.losses`,
`sot_b200.sharding`, `sot_b200.synthetic`, ...  No code lives here."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "1d-spectral-optimal-transport_b200")
__path__.append(_PKG_DIR)
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
