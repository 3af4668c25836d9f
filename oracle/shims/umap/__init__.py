"""Inert umap-learn stand-in (oracle/ test infrastructure); `vae_reg_GP.py:21`."""


class UMAP:
    def __init__(self, *a, **k):
        pass

    def fit_transform(self, x):
        raise RuntimeError("umap stub")
