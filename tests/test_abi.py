"""The drop-in boundary on CPU: libivgpt_b200.so loads without a GPU, exports every entry point include/ivgpt_b200.h
declares, the ctypes binding declares exactly the same set, and the descriptor structs have the C compiler's layout."""
import ctypes as C
import os
import re
import subprocess

import pytest

from helpers import ROOT

HEADER = os.path.join(ROOT, "include", "ivgpt_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ivgpt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ivideogpt_b200 import _lib
    names = _declared()
    assert len(names) >= 50
    lib = C.CDLL(_lib.LIB_PATH)                     # dlopen only: no CUDA call is made
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(_lib.SIGNATURES) == names, (sorted(set(_lib.SIGNATURES) ^ set(names)))
    handle = _lib.load()                            # the binding resolves every symbol with its signature
    assert handle.ivgpt_last_error() in (None, b"") or isinstance(handle.ivgpt_last_error(), bytes)
    assert int(handle.ivgpt_mega_layer_bytes()) == 6 * C.sizeof(C.c_void_p)
    assert int(handle.ivgpt_mega_packed_elems(17, 64)) == 32 * 64 and int(handle.ivgpt_mega_packed_elems_bn(17, 64, 48)) == 48 * 64
    assert int(handle.ivgpt_mega_packed_elems64(65, 128)) == 128 * 128


def test_descriptor_structs_match_the_c_layout(tmp_path):
    """sizeof / offsetof of the three descriptor structs as gcc sees the header vs the ctypes mirrors in _lib.py."""
    from ivideogpt_b200 import _lib
    prog = tmp_path / "layout.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "ivgpt_b200.h"\n'
        "int main(void) {\n"
        '  printf("%zu %zu %zu\\n", sizeof(ivgpt_gemm_desc), sizeof(ivgpt_conv_desc), sizeof(ivgpt_mega_desc));\n'
        '  printf("%zu %zu %zu %zu %zu\\n", offsetof(ivgpt_mega_desc, kcache), offsetof(ivgpt_mega_desc, prof),\n'
        "         offsetof(ivgpt_mega_desc, slot_emb), offsetof(ivgpt_mega_desc, qkvp), offsetof(ivgpt_mega_desc, bn_down));\n"
        "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    sizes, offs = [int(v) for v in out[:3]], [int(v) for v in out[3:]]
    assert sizes == [C.sizeof(_lib.GemmDesc), C.sizeof(_lib.ConvDesc), C.sizeof(_lib.MegaDesc)]
    M = _lib.MegaDesc
    assert offs == [M.kcache.offset, M.prof.offset, M.slot_emb.offset, M.qkvp.offset, M.bn_down.offset]


def test_product_never_imports_the_oracle_or_a_cpu_fallback():
    """The oracle is test infrastructure: nothing under ivideogpt_b200/ or ivideogpt/ may import oracle/ (or diffusers), and the
    library loader must raise -- not fall back -- when the .so is missing."""
    bad = []
    for base in ("ivideogpt_b200", "ivideogpt"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+(oracle|diffusers)\b", src, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    import importlib
    from ivideogpt_b200 import _lib
    saved_path, saved_handle = _lib.LIB_PATH, _lib._lib
    try:
        _lib.LIB_PATH, _lib._lib = os.path.join(ROOT, "ivideogpt_b200", "no_such_library.so"), None
        with pytest.raises(_lib.B200LibraryError, match="no CPU or PyTorch fallback"):
            _lib.load()
    finally:
        _lib.LIB_PATH, _lib._lib = saved_path, saved_handle
    importlib.import_module("ivideogpt_b200")       # importing the package never needs the GPU
