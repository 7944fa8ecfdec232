"""Stub of `matplotlib` (absent): imported at module scope by pMCTF/utils/util.py:10,16, never used on the hot path."""
