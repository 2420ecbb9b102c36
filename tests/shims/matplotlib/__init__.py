"""Import stub: the reference's CPU scripts `import matplotlib.pyplot` at module import
(gsplat/gausplat.py:1, backward_cpu.py:1) but this image has no matplotlib.  Only used by
tests/golden/make_golden.py and the reference-script runners (dev container only)."""
