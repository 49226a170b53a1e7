"""CPU: the C-ABI library builds, loads and exports every symbol include/geossl_b200.h declares
(no compute calls without a GPU), and the product refuses to run without CUDA."""
import os
import re

import pytest
import torch

from geossl_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(REPO, "include", "geossl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(geossl_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == _lib.exported_symbols()


def test_library_builds_and_exports_all_symbols():
    _lib.build()
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.geossl_abi_version() == _lib.ABI_VERSION == 4
    assert lib.geossl_filter_bwd_workspace(50, 128) > 0 and lib.geossl_ddm_workspace(128) > 0
    assert lib.geossl_filter_bwd_workspace(50, 48) == -1          # unsupported width is reported, not guessed


def test_sass_is_sm100a():
    _lib.build()
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    from geossl_b200 import ops
    from geossl_b200.Geom3D.models import SchNet
    pos = torch.rand(6, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.radius_csr(pos, torch.zeros(6, dtype=torch.long), 10.0, num_graphs=1)
    m = SchNet(hidden_channels=32, num_filters=32, num_interactions=1, num_gaussians=10, node_class=9)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(6, dtype=torch.long), pos, torch.zeros(6, dtype=torch.long))
    with pytest.raises(AssertionError):                              # schnet.py:86
        m(torch.zeros(6, dtype=torch.int32), pos)


def test_torch_ops_are_registered_and_cuda_only():
    """``torch.ops.geossl_b200.*`` (geossl_b200/torch_ops.py): every op has a schema, and a CPU tensor is refused by the
    dispatcher -- there is no CPU kernel to fall back to."""
    from geossl_b200 import torch_ops
    for name in torch_ops.OP_NAMES:
        assert hasattr(torch.ops.geossl_b200, name), name
    with pytest.raises(NotImplementedError, match="CPU"):
        torch.ops.geossl_b200.pair_distance(torch.zeros(3, 3), torch.zeros((2, 1), dtype=torch.long))
    with pytest.raises(NotImplementedError, match="CPU"):
        torch.ops.geossl_b200.cfconv(torch.zeros(2, 128), torch.zeros(1, 128), torch.zeros(3, dtype=torch.int32),
                                     torch.zeros(1, dtype=torch.int32))


def test_nvtx_switch_is_off_by_default_and_harmless():
    from geossl_b200 import ops
    assert ops.NVTX is False or os.environ.get("GEOSSL_NVTX")
    with ops.nvtx_range("noop"):
        pass
