"""Freeze oracle outputs into ``tests/golden/*.npz`` (TEST INFRASTRUCTURE, see ``oracle/__init__.py``).

Usage: ``python -m oracle.make_golden``.  The fixtures pin the restated oracle against regressions (the reference has
no golden vectors of its own -- parity unpinned); inputs are regenerated from the same seeds at test time.
"""
import os

import numpy as np

from .cases import CASES, ROOT, run_case


def main() -> None:
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in CASES:
        res = run_case(name)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **res)
        print(name, "->", path, {k: v.shape for k, v in list(res.items())[:6]})


if __name__ == "__main__":
    main()
