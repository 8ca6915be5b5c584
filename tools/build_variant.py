"""Build a named variant of the library next to the default one (A/B probes on the GPU box):

    python tools/build_variant.py timing -DDAG_TIMING        ->  magpy_rv_b200/_lib/variants/timing.so
    MRV_LIB=magpy_rv_b200/_lib/variants/timing.so python tools/dag_timing.py 1024x128

With a git revision as the first flag (``@<rev>``) the sources of that revision are compiled (baseline builds)."""
import os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpy_rv_b200 import build as b

name, flags = sys.argv[1], sys.argv[2:]
root = os.path.dirname(b.HERE)
out_dir = os.path.join(b.OUT_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, name + ".so")
src = b.SRC
tmp = None
if flags and flags[0].startswith("@"):
    rev = flags.pop(0)[1:]
    tmp = tempfile.mkdtemp()
    subprocess.run("git -C %s archive %s magpy_rv_b200/csrc include | tar -x -C %s" % (root, rev, tmp), shell=True, check=True)
    src = os.path.join(tmp, "magpy_rv_b200", "csrc", "abi.cu")
cmd = [b._nvcc()] + b.NVCC_FLAGS + flags + ["-o", out, src]
subprocess.run(cmd, check=True)
print(out)
