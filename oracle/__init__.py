"""ORACLE -- test infrastructure only.

CPU restatements of the reference hot path (and a ctypes driver for the unmodified reference
compiled into oracle/_ref/).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product package
slam_toolbox_b200 never does.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def build(verbose: bool = False) -> None:
    """Compile oracle/karto_port.c and, when /root/reference is present, the reference itself."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")
