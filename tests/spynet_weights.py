"""Seeded weights of ME_Spynet (reference key names and shapes, video_net.py:74-100), shared by oracle/make_golden.py and the tests
(1.44 M weights are regenerated instead of committed).  He-style scaling keeps the activations O(1) through the five 7x7 layers;
the last layer is scaled so a level adds a flow increment of the order of a pixel."""
import numpy as np

CH = ((8, 32), (32, 64), (64, 32), (32, 16), (16, 2))


def make(seed: int, L: int = 6):
    g = np.random.default_rng(seed)
    out = {}
    for lvl in range(L):
        for i, (ci, co) in enumerate(CH):
            std = (1.4 / np.sqrt(49 * ci)) if i < 4 else 0.3 / np.sqrt(49 * ci)
            out[f"moduleBasic.{lvl}.conv{i + 1}.weight"] = (std * g.standard_normal((co, ci, 7, 7))).astype(np.float32)
            out[f"moduleBasic.{lvl}.conv{i + 1}.bias"] = (0.05 * g.standard_normal(co)).astype(np.float32)
    return out
