"""The C-ABI shared library loads on a CPU-only host and exports every symbol include/rtucker.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(headers=("rtucker.h", "rtucker_debug.h")):
    """Entry points of the drop-in boundary (rtucker.h) and of the debug / self-test header (rtucker_debug.h)."""
    syms = set()
    for h in headers:
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        syms |= set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", text))
    return sorted(syms)


def test_debug_entry_points_are_not_in_the_public_header():
    public = declared_symbols(("rtucker.h",))
    for s in ("rt_tc_selftest", "rt_mma_probe", "rt_score_v3_set_profile", "rt_score_bce_v3_phases"):
        assert s not in public and s in declared_symbols(("rtucker_debug.h",))


def test_header_declares_the_path():
    syms = declared_symbols()
    for s in ("rt_rank_filtered", "rt_score_bce_fwd_bwd", "rt_query_fwd", "rt_query_bwd", "rt_gram", "rt_apply",
              "rt_small_retract", "rt_small_project", "rt_eigh", "rt_score_rank_fused"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    import rtucker_b200
    if not os.path.isfile(rtucker_b200.LIB_PATH):
        pytest.skip("library not built (run __graft_entry__.build())")
    handle = ctypes.CDLL(rtucker_b200.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, missing
    assert handle.rt_abi_version() == 1


def test_prototypes_cover_the_header():
    from rtucker_b200._lib import PROTOTYPES
    assert sorted(PROTOTYPES) == declared_symbols()


def test_no_cpu_fallback():
    """Ops refuse CPU tensors loudly instead of silently computing elsewhere."""
    import torch
    import rtucker_b200
    from rtucker_b200 import ops
    if not os.path.isfile(rtucker_b200.LIB_PATH):
        pytest.skip("library not built")
    with pytest.raises(rtucker_b200.RTuckerError):
        ops.gram(torch.zeros(4, 2), torch.zeros(4, 2))
