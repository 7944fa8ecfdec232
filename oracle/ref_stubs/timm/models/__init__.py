"""Stub of `timm` (absent in this image): the reference only needs trunc_normal_."""
