"""Extract DATA tables (not code) that parity requires verbatim from the reference.

    python tools/extract_tables.py

Writes dither_pie_b200/data/ostromoukhov_coeffs.npy: the 256 x 3 integer coefficient table of
OstromoukhovDitherStrategy.COEFFS_TABLE (dithering_lib.py:1170-1203).  The table is a published
constant (Ostromoukhov 2001) with reference-specific entries from index 93 on; it has to be
bit-identical for parity, so it is carried as data.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.ref_loader import load  # noqa: E402


def main():
    dl, _ = load()
    tab = np.asarray(dl.OstromoukhovDitherStrategy.COEFFS_TABLE, dtype=np.int16)
    assert tab.shape == (256, 3)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "dither_pie_b200", "data", "ostromoukhov_coeffs.npy")
    np.save(out, tab)
    print(out, tab.shape, tab.dtype)


if __name__ == "__main__":
    main()
