"""CPU oracle for the FruitNeRF / CropNeRF per-ray rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it, and only as the checker (or as the timed CPU baseline), never as a fallback for the CUDA path.

PARITY: WIRING PINNED TO THE REFERENCE'S OWN CODE, PRIMITIVES UNPINNED.  The reference (``/root/reference``) is a nerfstudio plugin; every arithmetic primitive on
the hot path is imported from nerfstudio 1.1.3 (pinned only by ``crop_nerf/Dockerfile:1``), which is neither
vendored in the reference tree nor installable in this image, and the reference ships no tests, golden
vectors or fixtures for this path.  This package therefore *restates* the published nerfstudio 1.1.3 torch
(``implementation="torch"``) algorithms and the reference's own wiring
(``crop_nerf/fruit_nerf/fruit_field.py``, ``fruit_nerf.py``, ``components/*.py``), each function citing the
lines it follows.

* Wiring (``oracle/fruit_torch.py``): pinned.  ``oracle/ref_shim.py`` imports the reference's ``fruit_field.py``,
  ``fruit_nerf.py`` and ``components/*.py`` unmodified from ``/root/reference`` and executes them with nerfstudio's
  primitives replaced by this restatement; the outputs, losses and gradient norms it produced are committed as
  ``tests/golden/ref_*.npz`` and ``tests/test_reference_pin_cpu.py`` requires the restated wiring to reproduce them bit
  for bit (train / eval / aabb / inference / export modes).
* Primitives (``oracle/nerfstudio_torch.py``: HashEncoding, MLP, SH, samplers, renderers, losses): parity unpinned --
  nerfstudio 1.1.3 is absent, so they are a restatement of its published torch algorithms checked only against
  hand-computed cases, torch autograd and -- where one exists in this image -- an implementation that is not ours or the
  integral the primitive is the closed form of (scipy's real spherical harmonics, the transmittance integral, numpy's
  piecewise-linear inverse cdf, the mip-NeRF 360 distortion double integral, a brute-force outer measure:
  ``tests/test_oracle_cpu.py``).  ``tests/golden/tiny_*.npz`` (``oracle/make_golden.py``) pin them against regressions,
  not against nerfstudio itself.
"""
