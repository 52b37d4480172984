"""The C-ABI library loads on a CPU-only box and exports every symbol include/vaecap.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        src = open(os.path.join(ROOT, "include", header)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names.update(re.findall(r"\b(vc_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_exported():
    from vae_captioning_b200 import build, lib
    build.build()
    h = lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, missing
    assert h.vc_abi_version() >= 1


def test_error_without_gpu_is_loud():
    """vc_create must fail with VC_E_CUDA (never fall back) when no B200 is present."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vae_captioning_b200 import lib
    from vae_captioning_b200.engine import Engine
    from vae_captioning_b200.parameters import Parameters
    p = Parameters()
    p.vocab_size = 100
    with pytest.raises(lib.VaecapError) as e:
        Engine(p, vocab_size=100, max_batch=2, max_len=4)
    assert "CUDA" in str(e.value) or "no CPU fallback" in str(e.value)


def test_config_struct_layout_matches_header():
    """ctypes mirror of vc_config has the header's field order (17 int32 then 10 float)."""
    from vae_captioning_b200.engine import VcConfig
    src = open(os.path.join(ROOT, "include", "vaecap.h")).read()
    body = re.search(r"typedef struct vc_config \{(.*?)\} vc_config;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(int32_t|float)\s+(\w+);", body)
    assert [n for _, n in fields] == [n for n, _ in VcConfig._fields_]
    for (ty, _), (_, cty) in zip(fields, VcConfig._fields_):
        assert cty is (ctypes.c_int32 if ty == "int32_t" else ctypes.c_float)


def test_comm_entry_points_without_gpu():
    """The data-parallel entry points bind libnccl lazily: drawing a unique id needs no GPU (0, or VC_E_NCCL when no
    libnccl can be loaded -- never a crash), and a null handle is an argument error, not a segfault."""
    from vae_captioning_b200 import lib
    h = lib.load()
    h.vc_comm_unique_id.argtypes = [ctypes.c_void_p]
    h.vc_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    buf = ctypes.create_string_buffer(128)
    rc = h.vc_comm_unique_id(buf)
    assert rc in (0, -4), (rc, h.vc_last_error())
    if rc == 0:
        buf2 = ctypes.create_string_buffer(128)
        assert h.vc_comm_unique_id(buf2) == 0 and buf.raw != buf2.raw  # ids are unique
    assert h.vc_comm_init(None, buf, 0, 1) == -1
    assert h.vc_comm_unique_id(None) == -1
