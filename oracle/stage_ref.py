"""Stage the UNMODIFIED reference package under ``oracle/_ref/`` so that it travels to the GPU box.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  ``/root/reference`` exists only in the build
container; ``oracle/_ref/`` is git-ignored (it never enters history) but not gpurun-ignored, so the GPU
box receives it with the snapshot, exactly like the built ``.so``.  The recipe installs the reference's
own wheel (``/root/reference/dist/magpy_rv-1.0.4-py3-none-any.whl``, whose modules are byte-identical to
``/root/reference/src/magpy_rv``) with ``pip install --no-deps --no-index --target oracle/_ref``; when pip
cannot do that it unpacks the wheel (a zip archive) itself.  Nothing is edited: matplotlib / seaborn,
which the reference imports at module top and which are not installed here, are stubbed at import time by
``oracle/ref_loader.py``.

    python oracle/stage_ref.py          # (re)stage; prints the staged path or why it could not
"""
import hashlib
import os
import shutil
import subprocess
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REF_ROOT = "/root/reference"
WHEEL = os.path.join(REF_ROOT, "dist", "magpy_rv-1.0.4-py3-none-any.whl")
MODULES = ("parameters", "kernels", "models", "auxiliary", "mcmc_aux", "gp_likelihood", "mcmc", "saving", "plotting")


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def staged():
    return os.path.isfile(os.path.join(DEST, "magpy_rv", "gp_likelihood.py"))


def stage(force=False):
    """Returns the staged directory, or None when the reference tree is not available here."""
    if staged() and not force:
        return DEST
    if not os.path.isfile(WHEEL):
        return None
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-deps", "--no-index", "--no-compile",
           "--target", DEST, WHEEL]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 or not staged():
        with zipfile.ZipFile(WHEEL) as z:
            z.extractall(DEST)
    # the wheel must be the same code as the source tree the survey cites (file:line)
    src = os.path.join(REF_ROOT, "src", "magpy_rv")
    for m in MODULES:
        a, b = os.path.join(src, m + ".py"), os.path.join(DEST, "magpy_rv", m + ".py")
        if os.path.isfile(a) and (not os.path.isfile(b) or _sha(a) != _sha(b)):
            raise RuntimeError("staged %s differs from %s" % (b, a))
    return DEST


if __name__ == "__main__":
    out = stage(force=True)
    print(out if out else "reference tree not found at %s: nothing staged" % REF_ROOT)
