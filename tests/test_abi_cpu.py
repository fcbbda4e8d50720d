"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/piranha_b200.h
declares, the ctypes structures match the header, and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import piranha_b200 as pb
from piranha_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "piranha_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = pb.load_library()
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert getattr(lib, n) is not None, n
    assert sorted(abi.EXPORTED_SYMBOLS) == names


def test_struct_layouts_match_the_header(tmp_path):
    """Compile the header with gcc and compare sizeof/offsetof of every struct field with the ctypes mirror."""
    structs = {"pk_poly": abi.PkPoly, "pk_result": abi.PkResult, "pk_opts": abi.PkOpts, "pk_stats": abi.PkStats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src_c = tmp_path / "layout.c"
    src_c.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src_c)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    src = open(HEADER).read()
    for code, name in enumerate(["PK_OK", "PK_E_ARGS", "PK_E_KEY_RANGE", "PK_E_CF_WIDTH", "PK_E_NOMEM", "PK_E_CUDA", "PK_E_DEGREE", "PK_E_UNSUPPORTED"]):
        assert re.search(rf"#define {name} {code}\b", src), name
        assert getattr(abi, name) == code
    for code, name in enumerate(["PK_ALGO_AUTO", "PK_ALGO_DENSE", "PK_ALGO_HASH", "PK_ALGO_ZONE", "PK_ALGO_BLOCK"]):
        assert re.search(rf"#define {name} {code}u\b", src), name
        assert getattr(abi, name) == code


def test_no_cpu_fallback():
    """Without a usable sm_100 device the context cannot be created: nothing computes on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the failure path is only reachable on a CPU box")
    with pytest.raises(pb.PkError):
        pb.Context(0)
    lib = pb.load_library()
    h = C.c_void_p()
    assert lib.pk_ctx_create(0, None, C.byref(h)) == abi.PK_E_CUDA
    assert not h.value
    # every entry point rejects a null context instead of computing anything
    r = abi.PkResult()
    assert lib.pk_mul(None, None, None, None, C.byref(r)) == abi.PK_E_ARGS


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "piranha_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"
