"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/sw4b200.h declares
(no compute calls without a GPU), and fails loudly - not silently - when no device is present."""
import os
import re

import pytest

from cudasw4_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sw4b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sw4_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTS) == declared
    assert lib.sw4_version().startswith(b"sw4b200")


def test_no_silent_fallback_without_gpu():
    import ctypes
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cudasw4_b200 as sw
    with pytest.raises(sw.SW4Error) as e:
        sw.CudaSW4()
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cudasw4_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "liboracle" not in src and "sw4o_" not in src and "oracle_lib" not in src, fn
    for fn in os.listdir(os.path.join(ROOT, "include")):
        assert "sw4o_" not in open(os.path.join(ROOT, "include", fn)).read()


def test_headers_compile_standalone(tmp_path):
    """include/sw4b200.h is plain C (no C++ / CUDA / torch types in the ABI); include/cudasw4.cuh, the source-compatible
    mirror of the reference's host class (src/cudasw4.cuh:244-2454), compiles with a host compiler alone."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "include")
    gcc, gxx = shutil.which("gcc"), "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gcc or not gxx:
        pytest.skip("no host compiler")
    c_src = tmp_path / "abi.c"
    c_src.write_text('#include "sw4b200.h"\nint main(void) { sw4_handle* h = 0; (void)h; return sw4_version() == 0; }\n')
    r = subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(c_src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cpp_src = tmp_path / "facade.cpp"
    cpp_src.write_text('#include "cudasw4.cuh"\n'
                       'int main() {\n'
                       '  cudasw4::KernelTypeConfig k; cudasw4::MemoryConfig m;\n'
                       '  cudasw4::CudaSW4 sw({0}, 10, cudasw4::BlosumType::BLOSUM62_20, k, m, false);\n'
                       '  sw.setGapOpenScore(-11); sw.setGapExtendScore(-1);\n'
                       '  cudasw4::ScanResult r = sw.scan("ACDE", 4); return (int)r.scores.size();\n'
                       '}\n')
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-I", inc, str(cpp_src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
