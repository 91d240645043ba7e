"""ORACLE -- test infrastructure only (see oracle/md_jk.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this package; the product package gpu4pyscf_b200 never does.
"""
from .oracle import get_jk, q_cond, build, fp   # noqa: F401
