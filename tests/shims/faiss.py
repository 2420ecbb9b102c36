"""Import stub for `faiss` (gsplat/read_write_model.py:41)."""
