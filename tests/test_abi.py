"""CPU: the C-ABI library loads and exports every symbol include/var_b200.h declares, the
ctypes prototypes cover the header one to one, and the product never routes through the oracle
or a CPU fallback.  No compute calls (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "var_b200.h")
PKG = os.path.join(ROOT, "voicecontrolledrobot-var_b200")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(var_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(vb):
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(vb._lib.lib, n), f"{n} declared in include/var_b200.h but not exported"


def test_ctypes_prototypes_cover_the_header(vb):
    assert sorted(vb._lib.PROTOTYPES) == header_functions()


def test_version_and_error_string(vb):
    assert vb._lib.lib.var_version() == 100
    assert isinstance(vb._lib.last_error(), str)


def test_net_plan_matches_reference_state_dict_layout(vb):
    """Tensor table of the C++ layer plan == reference state_dict keys / shapes / order (host only)."""
    import ctypes as C
    from oracle import model as omodel
    lib = vb._lib.lib
    for kind, net, F in ((0, omodel.KUKA, 100), (1, omodel.ITHOR, 600)):
        h = C.c_void_p()
        assert lib.var_net_create(kind, F, 3, C.byref(h)) == 0
        shapes = omodel.param_shapes(net)
        n = lib.var_net_num_tensors(h)
        assert n == len(shapes)
        name = C.create_string_buffer(128)
        nd, shp, off, pk = C.c_int(), (C.c_int * 4)(), C.c_int64(), C.c_int64()
        end = 0
        for i, (k, s) in enumerate(shapes.items()):
            assert lib.var_net_tensor_info(h, i, name, 128, C.byref(nd), shp, C.byref(off), C.byref(pk)) == 0
            assert name.value.decode() == k and tuple(shp[:nd.value]) == tuple(s)
            assert off.value == end and off.value % 4 == 0  # 16-byte aligned, densely packed
            end += pk.value
        assert end == lib.var_net_param_floats(h)
        n_ref = sum(int(__import__("numpy").prod(s)) for s in shapes.values())
        assert {omodel.KUKA: 213478, omodel.ITHOR: 3849126}[net] == n_ref <= end
        # workspace sizing is a dry run of the same allocator the launches use
        small = lib.var_net_workspace_bytes(h, 8, 16, 1)
        big = lib.var_net_workspace_bytes(h, 16, 32, 1)
        assert 0 < small < big
        assert lib.var_net_workspace_bytes(h, 8, 16, 0) < small
        lib.var_net_destroy(h)
    h = C.c_void_p()
    assert lib.var_net_create(0, 123, 3, C.byref(h)) == -3  # VAR_ERR_UNSUPPORTED: Kuka net is built for F=100


def test_product_never_imports_the_oracle_or_falls_back():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "import oracle" in src:
                    bad.append(f)
    assert not bad, bad


def test_missing_library_fails_loudly(tmp_path):
    import importlib.util
    import shutil
    dst = tmp_path / "pkg"
    dst.mkdir()
    shutil.copy(os.path.join(PKG, "_lib.py"), dst / "_lib.py")
    spec = importlib.util.spec_from_file_location("lonely_lib", dst / "_lib.py")
    mod = importlib.util.module_from_spec(spec)
    with pytest.raises(ImportError, match="no CPU fallback"):
        spec.loader.exec_module(mod)


def test_engine_refuses_to_run_without_cuda(vb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vb.VarEngine(vb.KUKA, 100, 3)


def test_host_gather_helpers(vb):
    """Host-side staging helpers of the streaming loader (no device access): row gather and ragged clip packing."""
    import torch
    lib = vb._lib.lib
    src = torch.arange(10 * 6, dtype=torch.uint8).reshape(10, 6)
    idx = torch.tensor([3, 0, 9, 9], dtype=torch.int64)
    dst = torch.zeros(4, 6, dtype=torch.uint8)
    assert lib.var_host_gather_rows(src.data_ptr(), 6, idx.data_ptr(), 4, dst.data_ptr(), 2) == 0
    assert torch.equal(dst, src[idx])
    arena = torch.arange(100, dtype=torch.int16)
    off = torch.tensor([10, -1, 40, 3], dtype=torch.int64)
    ln = torch.tensor([5, 0, 4, 7], dtype=torch.int64)
    out = torch.zeros(64, dtype=torch.int16)
    new_off = torch.zeros(4, dtype=torch.int64)
    n = lib.var_host_gather_clips(arena.data_ptr(), off.data_ptr(), ln.data_ptr(), 4, out.data_ptr(), new_off.data_ptr(), 3)
    assert n == 18 and new_off.tolist() == [0, -1, 6, 10]      # clips packed back to back, 4-byte aligned
    assert out[0:5].tolist() == [10, 11, 12, 13, 14] and out[6:10].tolist() == [40, 41, 42, 43]
    assert out[10:17].tolist() == list(range(3, 10))
