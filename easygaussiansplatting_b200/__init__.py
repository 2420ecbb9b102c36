"""B200-native (sm_100a) differentiable 3D Gaussian Splatting rasterizer: the hot path of
scomup/EasyGaussianSplatting (its `gsplatcu` extension) rebuilt from scratch.

  include/gsplat_b200.h            C ABI (the drop-in boundary)
  easygaussiansplatting_b200/csrc  CUDA kernels + C ABI -> libgsplat_b200.so
  easygaussiansplatting_b200.ops   the reference's seven-operator Python surface
  gsplatcu                         same operators under the reference's module name
"""
__version__ = "0.1.0"
