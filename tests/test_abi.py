"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every symbol include/nsb.h
declares, the Python binding covers them, and the product package never touches the oracle."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nsb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nsb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from nerf_sandbox_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/nsb.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "python binding and header disagree"
    lib = _lib.lib()
    assert lib.nsb_version() >= 100
    assert lib.nsb_error_string(-2).decode() == "workspace too small"
    assert lib.nsb_packed_weights_bytes() > 595844 * 4
    # sizes only -- no compute without a GPU
    assert lib.nsb_field_workspace_bytes(1024, 0, 1) > lib.nsb_field_workspace_bytes(1024, 0, 0) > 0
    assert lib.nsb_train_workspace_bytes(1024, 64, 128, 0) > 0


def test_product_never_imports_oracle_and_has_no_cpu_fallback():
    pkg = os.path.join(ROOT, "nerf_sandbox_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"
    import nerf_sandbox_b200 as nsb
    with pytest.raises(RuntimeError):
        nsb.get_vanilla_nerf_encoders()[0](torch.zeros(4, 3))          # CPU tensor -> loud failure
    with pytest.raises(RuntimeError):
        nsb.sample_pdf(torch.rand(2, 8), torch.rand(2, 8), 4, deterministic=True)
    m = nsb.NeRF(63, 27)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 63), torch.zeros(2, 27))
    with pytest.raises(RuntimeError):                                   # reference tests/unit/test_mlps.py:125-143
        m(torch.zeros(2, 60), torch.zeros(2, 27))


def test_module_interfaces_match_reference():
    import nerf_sandbox_b200 as nsb
    pe, de = nsb.get_vanilla_nerf_encoders()
    assert (pe.out_dim, de.out_dim) == (63, 27) and len(pe.state_dict()) == 0
    m = nsb.NeRF(63, 27, 8, 256, skip_pos=4)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes["mlp.0.weight"] == (256, 63) and shapes["mlp.4.weight"] == (256, 319)
    assert shapes["color_fc.weight"] == (128, 283) and shapes["sigma_out.weight"] == (1, 256)
    assert sum(p.numel() for p in m.parameters()) == 595844
    assert [l.in_features for l in m.mlp] == [63, 256, 256, 256, 319, 256, 256, 256]
    lines = []
    nsb.log_nerf_arch(m, logger=lines.append); m.enable_debug(3, lines.append); m._debug_dump_arch_once()
    assert any("SKIP" in l for l in lines)
    # flat storage aliases the parameters and survives load_state_dict
    flat = m.flat_params()
    sd = {k: torch.randn_like(v) for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    assert torch.equal(m.flat_params()[:256 * 63].view(256, 63), sd["mlp.0.weight"]) and m.flat_params() is flat
    with pytest.raises(ValueError):
        nsb.sample_pdf(torch.rand(2, 5), torch.rand(2, 8), 4)
    with pytest.raises(ValueError):
        nsb.sample_pdf(torch.rand(8), torch.rand(8), 4)
