"""Build ``lib/libb200va.so`` in-tree with nvcc for sm_100a (``python -m ...build``)."""

from __future__ import annotations

import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB = os.path.join(PKG_DIR, "lib", "libb200va.so")


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["make", "-C", CSRC, "-j", str(min(os.cpu_count() or 4, 8))]
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, stdout=subprocess.DEVNULL)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("building libb200va.so failed (nvcc -gencode arch=compute_100a,code=sm_100a)")
    if not os.path.exists(LIB):
        raise RuntimeError(f"{LIB} was not produced")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
