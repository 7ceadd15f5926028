"""The C-ABI library loads on a CPU-only box and exports every symbol include/dfw_b200.h declares."""
import ctypes
import os
import re

import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "dfw_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dfw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from deep_fem_uav_wing.gnn import _cabi

    declared = _declared_symbols()
    assert len(declared) >= 14
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/dfw_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == declared  # the ctypes table binds exactly the header
    assert _cabi.lib.dfw_abi_version() == 1


def test_size_queries_and_argument_errors_need_no_gpu():
    from deep_fem_uav_wing.gnn import _cabi

    lib = _cabi.lib
    assert lib.dfw_csr_ws_bytes(1000, 100) >= 8 * 1000 + 4 * 100
    assert lib.dfw_linear_bwd_weight_ws_bytes(1000, 128, 128, 128) > 0
    assert lib.dfw_epilogue_bwd_ws_bytes(1000, 128) > 0
    # argument validation happens before any CUDA call
    rc = lib.dfw_sage_aggregate(None, None, None, None, None, None, 10, 5, 3, 0, None)
    assert rc != 0 and b"multiple of 16" in lib.dfw_last_error()
    rc = lib.dfw_csr_build(None, -1, 5, 0, None, None, None, None, None, None, 0, None)
    assert rc != 0 and b"negative" in lib.dfw_last_error()
    with pytest.raises(_cabi.DfwError):
        _cabi.check(lib.dfw_linear_fwd(None, None, 0, None, None, 0, None, None, None, 1e-5, None, 0.0, 0, None, None, None,
                                       None, None, None, 4, 512, 0, 0, None, 0, None))


def test_one_call_entry_points_validate_before_any_cuda_call():
    from deep_fem_uav_wing.gnn import _cabi

    lib = _cabi.lib
    assert lib.dfw_sage_layer_fwd_ws_bytes(1000, 128, 128, 0) == lib.dfw_linear_ws_bytes(128, 128, 128, 0)
    with_gx, without = lib.dfw_sage_layer_bwd_ws_bytes(1000, 128, 128, 0, 1), lib.dfw_sage_layer_bwd_ws_bytes(1000, 128, 128, 0, 0)
    assert with_gx > without >= 1000 * 128 * 4  # g_y lives in the workspace, g_t only when an input gradient is wanted
    assert lib.dfw_mlp2_bwd_ws_bytes(1000, 10, 64, 128, 0, 0, 1) > lib.dfw_mlp2_bwd_ws_bytes(1000, 128, 64, 1, 0, 1, 1) > 0
    assert lib.dfw_sage_layer_fwd_ws_bytes(-1, 128, 128, 0) == 0 and lib.dfw_mlp2_fwd_ws_bytes(10, 0, 64, 1, 0) == 0
    rc = lib.dfw_sage_layer_fwd(None, None, None, None, None, None, None, None, None, 1e-5, 0.0, 0, 0, None, None, None, None,
                                10, 0, 128, 128, 0, None, 0, None)
    assert rc != 0 and b"dfw_sage_layer_fwd" in lib.dfw_last_error()
    rc = lib.dfw_mlp2_fwd(None, None, None, None, None, 0.0, 0, 0, 7, None, None, 10, 10, 64, 1, 0, None, 0, None)
    assert rc != 0 and b"unknown mode" in lib.dfw_last_error()
    rc = lib.dfw_csr_transpose(None, -1, 5, None, None, None, None, None, None, 0, None)
    assert rc != 0 and b"negative" in lib.dfw_last_error()


def test_product_path_refuses_cpu_tensors():
    import torch

    from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss, SAGEConv

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GraphSAGEModel(10, 16, 1, 1)(torch.randn(4, 10), torch.zeros(2, 0, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SAGEConv(8, 8)(torch.randn(4, 8), torch.zeros(2, 0, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MaskedMSELoss()(torch.randn(4, 1), torch.randn(4, 1), None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "deep-fem-uav-wing_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
