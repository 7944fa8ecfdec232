from torch.nn.init import trunc_normal_  # noqa: F401  (pMCTF_L.py:10, pWave.py:8)
