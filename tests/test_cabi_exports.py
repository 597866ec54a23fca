"""No-GPU checks of the drop-in boundary: the library loads, exports every symbol
include/dnet_b200.h declares, and fails loudly (never falls back) without a device."""
import ctypes as C

import pytest
import torch

from dnet_b200 import _cabi


def test_library_loads_and_exports_every_declared_symbol():
    lib = _cabi.load()
    declared = _cabi.declared_symbols()
    assert len(declared) >= 55
    missing = [s for s in declared if not hasattr(lib, s)]
    assert missing == []


def test_ctypes_prototypes_cover_the_header_exactly():
    assert set(_cabi._PROTOS) == set(_cabi.declared_symbols())


def test_header_has_no_torch_or_cxx_types():
    import re
    txt = _cabi.HEADER_PATH.read_text()
    assert 'extern "C"' in txt
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)   # declarations only, comments stripped
    for bad in ("torch", "at::", "std::", "c10"):
        assert bad not in txt
    assert _cabi.load().dn_version().startswith(b"dnet_b200")


def test_model_cfg_struct_layout_matches_header_order():
    import re
    txt = _cabi.HEADER_PATH.read_text()
    body = re.search(r"typedef struct dn_model_cfg \{(.*?)\} dn_model_cfg;", txt, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"(?:int32_t|float)\s+(\w+);", body)
    assert names == [f[0] for f in _cabi.ModelCfg._fields_]
    assert C.sizeof(_cabi.ModelCfg) == 4 * len(names)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device failure mode")
def test_no_gpu_fails_loudly_not_silently():
    lib = _cabi.load()
    rc = lib.dn_init(0)
    assert rc == _cabi.DN_ECUDA
    assert b"no CPU fallback" in lib.dn_last_error()
    with pytest.raises(_cabi.DnError):
        _cabi.init(0)
    from dnet_b200.shard.models import ShardLoadModelRequest
    from dnet_b200.shard.runtime import ShardRuntime

    rt = ShardRuntime("s")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rt.load_model_core(ShardLoadModelRequest(model_path="/nonexistent", total_layers=1, layers=[0], window_size=1,
                                                 residency_size=1, kv_bits="fp16"))


def test_product_never_imports_the_oracle():
    import pathlib
    root = pathlib.Path(_cabi.__file__).resolve().parent
    for p in root.rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
