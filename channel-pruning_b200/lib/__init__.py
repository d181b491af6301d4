"""Host-side mirror of the reference's ``lib`` package for the pruning hot path
(``from lib.decompose import *`` / ``from lib.net import Net`` in train.py:18-19)."""
from . import cfgs, decompose, net  # noqa: F401
