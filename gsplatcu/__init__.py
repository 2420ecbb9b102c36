"""Drop-in replacement for the reference's `gsplatcu` extension module.

`import gsplatcu as gsc` (reference gsplat/gsmodel.py:2, gsplat/utils.py:2,
forward_gpu.py:2, backward_gpu.py:3) resolves to this package when the repository root is on
sys.path (or after `pip install -e .`), so the reference's forward_gpu.py, backward_gpu.py,
train.py and gsplat/gsmodel.py run unmodified on the B200-native kernels.
The seven names are exactly the ones ext.cpp:68-76 registers.
"""
from easygaussiansplatting_b200.ops import (computeCov2D, computeCov3D, inverseCov2D, project,
                                            sh2Color, splat, splatB)

__all__ = ["splat", "splatB", "inverseCov2D", "computeCov3D", "computeCov2D", "project", "sh2Color"]
