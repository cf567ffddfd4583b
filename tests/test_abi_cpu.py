"""CPU: libofk.so loads and exports exactly the entry points include/ofk.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ofk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ofk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()  # nvcc cross-compiles for sm_100a without a GPU
    from open_flamingo_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = header_functions()
    assert declared, "no declarations parsed from include/ofk.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ofk.h but not exported by libofk.so"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes signature table and ofk.h disagree"
    loaded = _lib.lib()
    assert loaded.ofk_abi_version() == 2
    assert isinstance(_lib.launch_count(), int)


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU rather than silently computing on the CPU."""
    import pytest
    import torch
    from open_flamingo_b200 import ops
    from open_flamingo_b200.src.helpers import GatedCrossAttentionBlock, PerceiverResampler
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(16, 16, dtype=torch.bfloat16), torch.zeros(16, 16, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):
        PerceiverResampler(dim=64, depth=1)(torch.zeros(1, 1, 1, 4, 64))
    with pytest.raises(RuntimeError):
        GatedCrossAttentionBlock(dim=64, dim_visual=64)(torch.zeros(1, 4, 64), torch.zeros(1, 1, 64, 64),
                                                        media_locations=torch.zeros(1, 4, dtype=torch.bool))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "open_flamingo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), f"{f} references the oracle"
