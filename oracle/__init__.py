"""CPU oracle for the FruitNeRF / CropNeRF per-ray rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it, and only as the checker (or as the timed CPU baseline), never as a fallback for the CUDA path.

PARITY UNPINNED.  The reference (``/root/reference``) is a nerfstudio plugin; every arithmetic primitive on
the hot path is imported from nerfstudio 1.1.3 (pinned only by ``crop_nerf/Dockerfile:1``), which is neither
vendored in the reference tree nor installable in this image, and the reference ships no tests, golden
vectors or fixtures for this path.  This package therefore *restates* the published nerfstudio 1.1.3 torch
(``implementation="torch"``) algorithms and the reference's own wiring
(``crop_nerf/fruit_nerf/fruit_field.py``, ``fruit_nerf.py``, ``components/*.py``), each function citing the
lines it follows.  Golden vectors under ``tests/golden/`` are produced by *this* restatement from fixed
seeds (``oracle/make_golden.py``); they pin the oracle against regressions, not against nerfstudio itself.
"""
