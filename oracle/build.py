"""Build recipe for the C part of the oracle (TEST INFRASTRUCTURE, not product code).

    python -m oracle.build        -> oracle/_build/libdp_oracle.so

The reference (dobrosketchkun/dither_pie) is pure Python, so there is nothing to compile into
oracle/_ref; the oracle is a restatement ("port") whose parity is pinned by the golden vectors
under tests/golden/ that were generated from the live reference (tools/make_golden.py).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "dp_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libdp_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
