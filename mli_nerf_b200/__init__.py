"""mli_nerf_b200 -- B200 (sm_100a) implementation of MLI-NeRF's per-ray render hot path behind a C ABI.

  _lib      ctypes binding of libmli_b200.so (include/mli_b200.h); fails loudly if the library / device is missing
  build     nvcc recipe for the library
  engine    host-side orchestration of the kernels (HBM layout, call order)
  model     drop-in ``Model`` for ``--model.type=mli_nerf_b200.model``
  hashgrid  drop-in for ``tinycudann.Encoding`` (HashGrid)
  losses    fused in-kernel training losses
  dist      ray-sharded multi-GPU gradient exchange (NCCL)
"""
__version__ = "0.1.0"
