"""The C-ABI library loads on a CPU-only box and exports every symbol include/flowril_b200.h declares.
(No compute calls here: those need a GPU and live in the -m gpu tests.)"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from modril_b200 import build, _native
    build.build()
    return _native.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "flowril_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frl_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from modril_b200 import _native
    declared = _declared_symbols()
    assert declared, "no declarations found in the header"
    assert sorted(_native.EXPORTS) == declared
    raw = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} is declared in the header but not exported"


def test_abi_version_and_error_string(lib):
    from modril_b200 import _native
    assert lib.frl_abi_version() == _native.ABI_VERSION == 7
    assert isinstance(lib.frl_last_error(), bytes)


def test_invalid_arguments_return_status_not_crash(lib):
    from modril_b200 import _native
    # null config -> FRL_ERR_INVALID with a message; never touches the GPU
    h = ctypes.c_void_p()
    rc = lib.frl_net_create(None, ctypes.byref(h))
    assert rc == -1 and b"null" in lib.frl_last_error()
    assert lib.frl_net_destroy(None) == 0
    assert lib.frl_net_num_params(None) == -1
    with pytest.raises(ValueError):
        _native.check(rc, "frl_net_create")


def test_enums_match_header():
    from modril_b200 import _native
    text = open(os.path.join(ROOT, "include", "flowril_b200.h")).read()
    for name, val in (("FRL_PREC_FP32", _native.PREC_FP32), ("FRL_PREC_BF16", _native.PREC_BF16),
                      ("FRL_PREC_BF16X3", _native.PREC_BF16X3), ("FRL_PREC_FP16", _native.PREC_FP16),
                      ("FRL_PROBE_SHARED", 0), ("FRL_PROBE_PER_STEP", 1), ("FRL_PROBE_EXACT", 2),
                      ("FRL_REWARD_RAW", 0), ("FRL_REWARD_NORM", 1), ("FRL_REWARD_EXP", 2), ("FRL_REWARD_CLIP", 3),
                      ("FRL_REWARD_SMOOTH", 4)):
        m = re.search(name + r"\s*=\s*(-?\d+)", text)
        assert m and int(m.group(1)) == val, name


def test_header_is_plain_c(tmp_path, lib):
    """The boundary is a C ABI: the header compiles as C99 (-pedantic) and a C translation unit that takes the address of
    every declared entry point links against the library (no C++ or torch types in any signature)."""
    import shutil
    import subprocess

    from modril_b200 import _native

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = os.path.join(root, "include", "flowril_b200.h")
    subprocess.check_call(["gcc", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", header])
    names = _declared_symbols()
    assert set(names) == set(_native.EXPORTS)
    src = tmp_path / "link_all.c"
    body = "\n".join(f"    p[{i}] = (void (*)(void))&{n};" for i, n in enumerate(names))
    src.write_text('#include "flowril_b200.h"\n#include <stdio.h>\nint main(void) {\n    void (*p[%d])(void);\n%s\n'
                   '    printf("%%d %%d\\n", (int)(p[0] != 0), %d);\n    return 0;\n}\n' % (len(names), body, len(names)))
    exe = tmp_path / "link_all"
    lib_dir = os.path.dirname(_native.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wno-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                           "-L", lib_dir, "-lflowril_b200", f"-Wl,-rpath,{lib_dir}"])
