"""The C-ABI library loads and exports every symbol include/rtd3.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rtd3.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtd3_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "rtd3_env_step" in syms and "rtd3_version" in syms


def test_library_exports_every_declared_symbol(pkg):
    path = pkg._lib.LIB_PATH
    assert os.path.isfile(path), "librtd3.so not built: run __graft_entry__.build()"
    L = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_binding_table_matches_header(pkg):
    assert pkg._lib.exported_symbols() == declared_symbols()


def test_version_and_error_string(pkg):
    L = pkg._lib.lib()
    assert L.rtd3_version() == 100
    assert isinstance(L.rtd3_last_error(), bytes)


def test_argument_errors_are_reported_without_a_gpu(pkg):
    L = pkg._lib.lib()
    rc = L.rtd3_env_step(None, None, None, None, None, 4, 0, None)
    assert rc == -1
    assert b"rtd3_env_step" in L.rtd3_last_error() or b"launch_step" in L.rtd3_last_error()
    with pytest.raises(pkg._lib.Rtd3Error):
        pkg._lib.check(rc, "env_step")


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        pkg.Environment()


def test_product_does_not_import_oracle():
    pkg_dir = os.path.join(ROOT, "residual-td3-robot-navigation_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_tick_state_struct_layout_matches_the_header(pkg, tmp_path):
    """`_lib.TickStateStruct` (ctypes) against `rtd3_tick_state` as gcc lays it out from include/rtd3.h: every field offset."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    T = pkg._lib.TickStateStruct
    names = [f[0] for f in T._fields_]
    src = tmp_path / "off.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rtd3.h"\nint main(void){\n'
                   + "".join('printf("%%zu\\n", offsetof(rtd3_tick_state, %s));\n' % k for k in names)
                   + 'printf("%zu\\n", sizeof(rtd3_tick_state));\nreturn 0;}\n')
    exe = tmp_path / "off"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [getattr(T, k).offset for k in names] + [ctypes.sizeof(T)]


def test_new_entry_points_reject_bad_arguments_without_a_gpu(pkg):
    """Argument checks of the fused-tick / candidate-list / f16-forward / peer-memory entry points run before any launch."""
    L, lib = pkg._lib.lib(), pkg._lib
    t = lib.TickStateStruct()
    assert L.rtd3_tick_pre(None, None) == -1
    assert L.rtd3_tick_pre(ctypes.byref(t), None) == -1 and b"null" in L.rtd3_last_error()           # every pointer of the struct is NULL
    assert L.rtd3_tick_post(None, ctypes.byref(t), None, None, 0, None) == -1
    assert L.rtd3_tick_run_f16(None, ctypes.byref(t), 256, 2, None, None, 0, 8, 0, None) == -1
    assert L.rtd3_demo_lists(None, 10, None, None, None, None) == -1
    one = ctypes.c_double(0.0)
    assert L.rtd3_demo_lists(ctypes.byref(one), 0, None, None, None, None) == -1                        # an empty set has no lists
    assert L.rtd3_demo_lists(ctypes.byref(one), 5, None, None, None, None) == -1                        # count pass without counts
    assert L.rtd3_mlp_forward_f16(200, 2, 0, ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), 128, None) == -1
    assert L.rtd3_mlp_forward_f16(256, 3, 0, ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), 128, None) == -1
    assert L.rtd3_mlp_forward_f16(256, 2, 6, ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), 128, None) == -1
    ptrs = (ctypes.c_void_p * 2)(0, 0)
    assert L.rtd3_p2p_allreduce(ptrs, ptrs, 0, 1, ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), 16, 16, ctypes.byref(one), None) == -1   # world 1
    assert L.rtd3_p2p_allreduce(ptrs, ptrs, 0, 9, ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), 16, 16, ctypes.byref(one), None) == -1   # world 9
    assert L.rtd3_p2p_allreduce(ptrs, ptrs, 0, 2, ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), 6, 16, ctypes.byref(one), None) == -1    # count % 4
    assert L.rtd3_p2p_allreduce(ptrs, ptrs, 0, 2, ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), 16, 16, ctypes.byref(one), None) == -1   # null peers
    assert L.rtd3_p2p_allreduce(ptrs, ptrs, 0, 2, ctypes.byref(one), ctypes.byref(one), ctypes.byref(one), 16, 8, ctypes.byref(one), None) == -1    # slot < count


def test_round2_entry_points_reject_bad_arguments_without_a_gpu(pkg):
    """rtd3_td3_update / rtd3_comm_* / rtd3_allreduce_grads / rtd3_td3_target_noise check their arguments before touching a device."""
    L, lib = pkg._lib.lib(), pkg._lib
    a = lib.Td3UpdateArgs()
    assert L.rtd3_td3_update(None, ctypes.byref(a), None) == -1
    assert L.rtd3_allreduce_grads(None, None, 4, None) == -1
    assert L.rtd3_comm_unique_id(None) == -1
    assert L.rtd3_comm_create(None, None, 0, 1, 0) == -1
    assert L.rtd3_comm_world(None) == -1 and L.rtd3_comm_rank(None) == -1 and L.rtd3_comm_destroy(None) == 0
    assert L.rtd3_td3_target_noise(1, 2, None, 8, None) == -1
    assert L.rtd3_comm_nccl_version() >= 20000          # resolved at run time from the NCCL copy torch ships


def test_ctypes_structs_match_the_header_layout(pkg, tmp_path):
    """The HOST structs of the ABI as ctypes sees them == as a C compiler lays them out from include/rtd3.h (sizeof + every offset)."""
    import subprocess
    lib = pkg._lib
    structs = {"rtd3_tick_state": lib.TickStateStruct, "rtd3_td3_update_args": lib.Td3UpdateArgs, "rtd3_p2p_state": lib.P2pStateStruct,
               "rtd3_mt_bank": lib.MtBankStruct}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rtd3.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for field, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, field, cname, field))
    lines.append("return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for field, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, field)]) == getattr(cls, field).offset, (cname, field)
