"""Recipe: stage the UNMODIFIED reference (meijieru/AtomNAS) under oracle/_ref/ for the GPU box.

TEST INFRASTRUCTURE.  The reference is pure Python; "building" it is copying its own files, where
they lie under /root/reference, into oracle/_ref/AtomNAS/ (git-ignored, NOT gpurun-ignored, so it
travels to the GPU box exactly like a built .so).  Nothing is edited.  Two tiny stubs for
third-party packages that are absent from this image are written NEXT to it (oracle/_ref/stubs/):
`lmdb` (imported by utils/lmdb_dataset.py:8, never called with --dataset imagenet1k_fake) and a
`tensorboard`-free `SummaryWriter` (common.py:9).

    python oracle/make_ref.py            # no-op (exit 0) when /root/reference is not present

Used by: bench.py --impl reference / reference-gpu and cpu_baseline (kind "reference"),
tests that compare against the real reference on the GPU box.  Never imported by atomnas_b200/.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ATOMNAS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

LMDB_STUB = '''"""Stub of the `lmdb` package (absent from this image).  The synthetic-data path
(--dataset imagenet1k_fake, utils/dataflow.py:185-192) never opens a database."""


def open(*args, **kwargs):
    raise RuntimeError("lmdb is not installed in this image; use --dataset imagenet1k_fake")
'''

TB_STUB = '''"""No-op stand-in for torch.utils.tensorboard.SummaryWriter (the `tensorboard` package is absent
from this image; common.py:9 imports it at module level)."""


class SummaryWriter(object):
    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, name):
        def _noop(*args, **kwargs):
            return None
        return _noop
'''


def main():
    if not os.path.isdir(SRC):
        print("make_ref: %s not present; keeping whatever is in %s" % (SRC, DST))
        return 0
    tree = os.path.join(DST, "AtomNAS")
    if os.path.isdir(tree):
        shutil.rmtree(tree)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(SRC, tree, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc", ".SUBMODULES.json"))
    stubs = os.path.join(DST, "stubs")
    os.makedirs(stubs, exist_ok=True)
    with open(os.path.join(stubs, "lmdb.py"), "w") as f:
        f.write(LMDB_STUB)
    with open(os.path.join(stubs, "tb_stub.py"), "w") as f:
        f.write(TB_STUB)
    n = sum(len(fs) for _, _, fs in os.walk(tree))
    print("make_ref: staged %d files of %s under %s" % (n, SRC, tree))
    return 0


if __name__ == "__main__":
    sys.exit(main())
