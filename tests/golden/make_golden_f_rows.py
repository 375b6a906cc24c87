"""Generates tests/golden/f_rows_small.npz: the oracle's checksums of the SURVEY.md 8(f) rows (occupancy, colour, decay with
deallocation, freespace, 2-D ESDF slice, markUnobservedTsdfFreeInsideRadius) on the inputs of c2_small.npz. The reference itself
cannot be built or imported here, so the vectors come from the oracle that is pinned to the reference's tests.
Run from the repo root:  python tests/golden/make_golden_f_rows.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_f_rows import run_oracle  # noqa: E402


def main():
    g = np.load(os.path.join(HERE, "c2_small.npz"))
    out = run_oracle(g)
    np.savez_compressed(os.path.join(HERE, "f_rows_small.npz"), **{k: np.int64(v) for k, v in out.items()})
    print(out)


if __name__ == "__main__":
    main()
