"""The C-ABI shared library loads on a machine without a GPU and exports every symbol that
include/stb.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from scannertools_b200 import _lib


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'stb.h')).read()
    return sorted(set(re.findall(r'STB_API\s+[^;{]*?\b(stb_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def built_lib():
    from scannertools_b200 import build
    build.build()
    return _lib.load()


def test_header_and_binding_table_agree():
    assert declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(built_lib):
    for name in declared_symbols():
        assert hasattr(built_lib, name), name


def test_no_device_calls_fail_loudly_not_silently(built_lib):
    assert built_lib.stb_version() >= 100
    p = _lib.FarnebackParams()
    built_lib.stb_farneback_default_params(C.byref(p))
    assert (p.num_levels, p.pyr_scale, p.win_size, p.num_iters, p.poly_n, p.poly_sigma, p.flags) == (3, 0.5, 15, 3, 5, 1.2, 0)
    # 1080p, 16 pairs: a few GB of workspace, well inside 180 GB
    ws = built_lib.stb_farneback_workspace_bytes(1920, 1080, 16, None)
    assert 1 << 30 < ws < 8 << 30
    assert built_lib.stb_farneback_workspace_bytes(0, 1080, 16, None) == 0
    bad = _lib.FarnebackParams(3, 0.3, 0, 15, 3, 5, 1.2, 0)          # pyr_scale < 0.5 is not implemented
    assert built_lib.stb_farneback_workspace_bytes(640, 480, 1, C.byref(bad)) == 0
    if built_lib.stb_device_count() == 0:
        h = C.c_void_p()
        rc = built_lib.stb_farneback_create(64, 64, 1, None, C.byref(h))
        assert rc != 0 and not h.value
        assert built_lib.stb_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'scannertools_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.h', '.cpp', '.cuh')):
                txt = open(os.path.join(dirpath, f), errors='replace').read()
                assert 'import oracle' not in txt and 'from oracle' not in txt and 'cuda_emu.h"' not in txt.replace('#include "cuda_emu.h"\n#else', ''), f


def test_header_is_plain_c():
    """The boundary is a C ABI: include/stb.h must compile as C99 on its own (no C++, no CUDA or
    torch types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    hdr = os.path.join(ROOT, 'include', 'stb.h')
    r = subprocess.run([gcc, '-x', 'c', '-std=c99', '-fsyntax-only', '-Wall', '-Werror', hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = re.sub(r'/\*.*?\*/', '', open(hdr).read(), flags=re.S)       # declarations only, comments stripped
    assert 'torch' not in code and 'cudaStream_t' not in code and 'std::' not in code and '#include <cuda' not in code
