"""CPU oracle for the DDPM / PixelCNN / VQ hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline.  The product path (``image-generation-models_b200``) never
imports this package and fails loudly when its CUDA library is missing.
"""
