"""Import the UNMODIFIED reference for pinning the oracle and for `bench.py --impl reference`.

TEST / BENCH INFRASTRUCTURE ONLY -- the product (`sot_b200`) never imports this.

Two places can hold the reference's own files:
  * `/root/reference`            -- the build container only;
  * `<repo>/baseline/_ref/`      -- a verbatim, git-ignored copy of the few pure-Python files the hot path's
                                    callers need (`stage()` below makes it when `/root/reference` is present;
                                    `__graft_entry__.build()` calls it).  It is not part of the repository's
                                    history but travels with the snapshot to the GPU box, where `/root/reference`
                                    does not exist -- so the reference arm there is the reference's stock code.
`features.py:7,10` imports nnAudio and librosa at module top (absent in this image, unused by `losses.py`), so
empty stand-in modules are registered before the import; no reference file is modified.
"""
import filecmp
import os
import shutil
import sys
import types

REFERENCE_ROOT = "/root/reference"
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(REPO_ROOT, "baseline", "_ref")
# losses.py imports features + utils; the training-step example (row f2) drives encoder / synths / ddsp
FILES = ("losses.py", "utils.py", "features.py", "encoder.py", "synths.py", "ddsp.py", "LICENSE")


def root():
    """Directory that holds the reference's files, or None."""
    for r in (REFERENCE_ROOT, STAGED_ROOT):
        if os.path.isfile(os.path.join(r, "losses.py")):
            return r
    return None


def available() -> bool:
    return root() is not None


def stage() -> bool:
    """Verbatim copy of FILES from /root/reference into baseline/_ref/ (git-ignored).  Returns True if the staged
    copy exists afterwards."""
    if os.path.isfile(os.path.join(REFERENCE_ROOT, "losses.py")):
        os.makedirs(STAGED_ROOT, exist_ok=True)
        for name in FILES:
            src, dst = os.path.join(REFERENCE_ROOT, name), os.path.join(STAGED_ROOT, name)
            if os.path.isfile(src) and not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
                shutil.copyfile(src, dst)
    return os.path.isfile(os.path.join(STAGED_ROOT, "losses.py"))


def _stub_absent_dependencies():
    for name in ("nnAudio", "nnAudio.features", "librosa"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["nnAudio"].features = sys.modules["nnAudio.features"]


def load(name: str = "losses"):
    """Returns the reference's module `name` (default `losses`; makes `features`, `utils`, ... importable)."""
    r = root()
    if r is None:
        raise RuntimeError("reference files not present (neither /root/reference nor baseline/_ref/)")
    _stub_absent_dependencies()
    if r not in sys.path:
        sys.path.append(r)
    mod = sys.modules.get(name)
    if mod is not None and not getattr(mod, "__file__", "").startswith(r):
        raise RuntimeError(f"a different module named {name!r} is already imported")
    return __import__(name)
