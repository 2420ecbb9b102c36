#!/usr/bin/env python
"""gau_loss (row N2) at 1920x1080: per-kernel CUDA-event times; also the ncu target:
  ncu --set full --clock-control none -k regex:k_ssim -s 2 -c 2 -o prof_loss python benchmarks/profile_loss.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easygaussiansplatting_b200 import _lib  # noqa: E402
from easygaussiansplatting_b200.loss import gau_loss_with_grad  # noqa: E402

H, W = 1080, 1920
img = torch.rand((3, H, W), device="cuda")
gt = (img + 0.05 * torch.randn_like(img)).clamp(0, 1)
lib = _lib.load()
for _ in range(3):
    gau_loss_with_grad(img, gt)
torch.cuda.synchronize()
lib.gsb_profile_enable(1)
for _ in range(10):
    gau_loss_with_grad(img, gt)
torch.cuda.synchronize()
lib.gsb_profile_enable(0)
for i in range(lib.gsb_profile_kernels()):
    tot, cnt = C.c_double(0), C.c_longlong(0)
    lib.gsb_profile_read(i, C.byref(tot), C.byref(cnt))
    if cnt.value:
        print(lib.gsb_profile_kernel_name(i).decode(), "%.4f ms" % (tot.value / cnt.value))
