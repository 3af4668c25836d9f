"""Inert stand-in for matplotlib (absent from this image).

TEST INFRASTRUCTURE ONLY (oracle/): lets the read-only reference import
(`vae_reg_GP.py:11-12`, `utils.py:14-15`).  Every attribute is a no-op.
"""


class _Inert:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return self


def __getattr__(name):
    return _Inert()
