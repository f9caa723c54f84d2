"""B200-native (sm_100a) differentiable video-Gaussian rasterizer.

Layout:
  csrc/      hand-written CUDA kernels + the C ABI (include/spv_b200.h) -> libspv_b200.so
  gs/        drop-in mirror of the reference's ``dptr.gs`` operator module (autograd Functions over the C ABI)
  renderer/  mirrors of the pointrix renderer classes that sit on top of ``dptr.gs``
  synth.py   synthetic DAVIS-shaped scenes (the only data source of the bench and the tests)
"""
__version__ = "0.1.0"
