"""Import the UNMODIFIED reference (`/root/reference`) for pinning the oracle -- build container only.

TEST INFRASTRUCTURE ONLY.  `/root/reference` does not exist on the GPU box; nothing that runs
there (`-m gpu` tests, smoke(), bench.py) may call this.  `features.py:7,10` imports nnAudio and
librosa at module top (absent here, unused by `losses.py`), so two empty stand-in modules are
registered before the import; no reference file is modified or copied.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "losses.py"))


def load():
    """Returns the reference's `losses` module (and makes `features`, `utils` importable)."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    for name in ("nnAudio", "nnAudio.features", "librosa"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["nnAudio"].features = sys.modules["nnAudio.features"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    mod = sys.modules.get("losses")
    if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("a different module named 'losses' is already imported")
    import losses  # noqa: the reference module

    return losses
