"""ORACLE -- test infrastructure only (see lfo_oracle.c header).

A CPU restatement of the linfa-linalg Householder/Cholesky/triangular hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product
(linfa_linalg_b200) never does.
"""
from .ref import *  # noqa: F401,F403
