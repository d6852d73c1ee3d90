"""aon_b200 -- B200-native (sm_100a) volume-rendering hot path of zubair-irshad/articulated-object-nerf.

Package directory ``articulated-object-nerf_b200/`` (import it as ``aon_b200``):

* ``csrc/``      hand-written CUDA kernels + the C ABI of ``include/aon.h`` -> ``libaon_b200.so``
* ``lib.py``     ctypes binding (fails loudly if the library is missing; no CPU fallback)
* ``nerf.py``    ``NeRF`` / ``NeRF_AE_Art`` / ``CodeLibraryArticulated`` with the reference's signatures
* ``lit.py``     ``LitNeRF`` / ``LitNeRF_AutoDecoder`` hook surface + a minimal trainer (no Lightning in the image)
* ``dist.py``    ray-sharded multi-GPU render (one process per GPU, one all-gather per image)
"""
__version__ = "0.1.0"
