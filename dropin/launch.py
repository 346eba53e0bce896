"""Run one of the reference's entry scripts (train.py / val.py) on the B200 engine, unchanged.

    torchrun --nproc_per_node=8 /path/to/atomnas-b200/dropin/launch.py \
        /path/to/AtomNAS/train.py app:apps/slimming/shrink/atomnas_c.yml --dataset imagenet1k_fake

Python puts the directory of the script it runs FIRST on sys.path, so `python AtomNAS/train.py`
would always import the reference's own `models/` and `utils/` whatever PYTHONPATH says.  This
launcher orders the path explicitly -- dropin/ (the shims), this repo, then the reference
checkout -- and executes the script as __main__ with its own argv.  The reference checkout is the
script's directory unless ATOMNAS_REFERENCE names it.
"""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        sys.exit("usage: launch.py <reference script> [its arguments...]")
    script = os.path.abspath(sys.argv[1])
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    ref = os.path.abspath(os.environ.get("ATOMNAS_REFERENCE") or os.path.dirname(script))
    os.environ["ATOMNAS_REFERENCE"] = ref
    head = [here, repo, ref]
    sys.path[:] = head + [p for p in sys.path if os.path.abspath(p or os.curdir) not in head and p != ""]
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
