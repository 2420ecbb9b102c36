"""End-to-end use of the gradients: a short Adam run against images of a hidden scene must
reduce the L1 loss substantially on both the fused path and the seven-operator surface, and
the two must follow the same loss curve (BASELINE config 3 stand-in, see
benchmarks/train_synthetic.py)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "benchmarks"))


def test_training_reduces_loss_and_paths_agree():
    from train_synthetic import train
    f1, l1, c1 = train(n=4000, views=2, iters=80, W=160, H=120, use_ops=False, verbose=False)
    f2, l2, c2 = train(n=4000, views=2, iters=80, W=160, H=120, use_ops=True, verbose=False)
    assert l1 < 0.55 * f1, (f1, l1)
    assert l2 < 0.55 * f2, (f2, l2)
    assert abs(f1 - f2) < 1e-5
    # same optimisation trajectory up to fp32 noise amplified over 80 Adam steps
    assert abs(l1 - l2) < 0.05 * l1, (l1, l2)
    assert max(abs(a - b) for a, b in zip(c1[:10], c2[:10])) < 1e-4
