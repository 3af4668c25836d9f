"""Inert matplotlib.pyplot stand-in (see package docstring)."""
from . import _Inert

_inert = _Inert()


def subplots(*a, **k):
    return _inert, _inert


def __getattr__(name):
    return _inert
