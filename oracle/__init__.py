"""CPU oracle for the maua audio-reactive render path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker.  The product
(``maua_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * oracle.sg2       pinned against the reference's in-tree inference network
                     (``maua/GAN/wrappers/inference``), fixtures in tests/golden.
  * oracle.audio     pinned against the reference's torch-native features
                     (``maua/audiovisual/audioreactive/selfsupervised/features``).
  * oracle.signal    pinned against ``maua/audiovisual/audioreactive/signal.py``.
  * oracle.noise     pinned against ``selfsupervised/noise.py`` (the reference's own classes).
  * oracle.image     pinned against ``maua/ops/image.py`` resample.
  * oracle.selfsup   pinned against ``selfsupervised/features/processing.py`` and
                     ``selfsupervised/latent.py`` (tests/golden/make_selfsup_golden.py);
                     its natural cubic spline stands in for the absent
                     torchcubicspline (restated, checked against scipy: unpinned).
  * oracle.warp      PARITY UNPINNED: kornia (absent, un-pinned) translate / rotate /
                     scale restated on top of torch's affine_grid / grid_sample.
  * oracle.sg3       PARITY UNPINNED: the StyleGAN3 network lives in the
                     un-vendored submodule maua/GAN/nv (maua-maua-maua/nvGAN @
                     7809c05, a fork of NVlabs/stylegan3); the restatement
                     follows the published upstream algorithm and is anchored on
                     the reference's own call sites and ``layer_multipliers``.
"""
