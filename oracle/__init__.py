"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MLI-NeRF (NeuralLumen) per-ray render path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py`` (its ``cpu_baseline`` leg and ``--impl reference``) may import it, and only as the checker /
the CPU baseline -- never as the thing shipped.  The product (``mli_nerf_b200``) raises if its CUDA
library is missing; it has no CPU fallback and never imports this package.

Contents
  torch_hashgrid.py  fp32 torch restatement of tiny-cuda-nn's HashGrid encoding (the one piece of the
                     reference's arithmetic that is NOT under /root/reference).  PARITY UNPINNED: tcnn is
                     an un-vendored, un-pinned dependency and the reference has no tests at that boundary.
  port.py            torch-CPU restatement of the whole hot path (travels to the GPU box).
  ref_import.py      imports the UNMODIFIED reference from /root/reference (this container only) to pin
                     port.py and to generate tests/golden/*.
  gen_golden.py      the committed script that produced tests/golden/*.
  c/                 plain-C restatement of the integer parts (hash-grid corner index, inverse-CDF bins).
"""
