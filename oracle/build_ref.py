"""Recipe that installs the UNMODIFIED reference package into oracle/_ref/ — TEST/BENCH INFRASTRUCTURE.

    python oracle/build_ref.py            # only works where /root/reference exists (the build container)

The reference (lrgr/munk) is a pure-Python package (munk/setup.py:25-31: no extension modules), so
"building" it is `pip install --no-deps --target oracle/_ref` of `/root/reference/munk` from a
scratch copy (pip writes egg-info next to setup.py and /root/reference is read-only).  Nothing is
copied into the repository's history: oracle/_ref/ is git-ignored, but it is NOT gpurun-ignored, so
the installed package travels to the GPU box, where /root/reference does not exist.

Users: bench.py (`--impl reference` and the `cpu_baseline` leg, kind "reference") and
tests/test_oracle.py (pins the oracle restatement against it when present).  `load()` imports it
under the in-memory `sklearn.externals.joblib` alias the 2018 sources need;
no reference file is edited.
"""
import importlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SOURCE = "/root/reference/munk"


def installed():
    return os.path.exists(os.path.join(REF_DIR, "munk", "munk.py"))


def build(force=False):
    """pip-installs the reference into oracle/_ref; returns True if it is present afterwards."""
    if installed() and not force:
        return True
    if not os.path.isdir(SOURCE):
        return installed()
    tmp = tempfile.mkdtemp(prefix="munk_ref_src_")
    try:
        src = os.path.join(tmp, "munk")
        shutil.copytree(SOURCE, src)
        if os.path.isdir(REF_DIR):
            shutil.rmtree(REF_DIR)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
               "--no-deps", "--no-compile", "--target", REF_DIR, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            return False
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return installed()


def load():
    """The reference `munk` package (module object) imported from oracle/_ref, or None.

    The repository root ships its own `munk` alias package (the drop-in); to import the reference
    under the same name the alias is kept out of the way: oracle/_ref goes first on sys.path and any
    `munk*` module already imported from elsewhere is refused rather than silently reused."""
    if not installed():
        return None
    import joblib
    import sklearn
    if "sklearn.externals.joblib" not in sys.modules:
        ext = sys.modules.get("sklearn.externals") or type(sys)("sklearn.externals")
        ext.joblib = joblib
        sklearn.externals = ext
        sys.modules["sklearn.externals"] = ext
        sys.modules["sklearn.externals.joblib"] = joblib
    have = sys.modules.get("munk")
    if have is not None:
        path = os.path.abspath(getattr(have, "__file__", "") or "")
        if not path.startswith(REF_DIR + os.sep):
            raise ImportError("a different `munk` package is already imported from %s" % path)
        return have
    sys.path.insert(0, REF_DIR)
    try:
        mod = importlib.import_module("munk")
    finally:
        sys.path.remove(REF_DIR)
    assert os.path.abspath(mod.__file__).startswith(REF_DIR + os.sep), mod.__file__
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "installed" if ok else "NOT installed (no /root/reference here)")
    sys.exit(0 if ok else 1)
