"""The C ABI: every symbol include/npt.h declares is exported by the built library, the ctypes mirror matches the C
struct layout, and the product fails loudly (NPT_ENODEVICE) without a GPU instead of falling back to the CPU."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from nptranscript_b200 import _build, abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "npt.h")


def declared_symbols():
    txt = open(HDR).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(npt_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_match_python_list():
    assert declared_symbols() == sorted(abi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_build.build_libnpt())
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert lib.npt_abi_version() == abi.NPT_ABI_VERSION


def test_struct_layout_matches_c():
    names = ["npt_config", "npt_annotation", "npt_batch", "npt_results", "npt_device_view"]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "npt.h"\nint main(){\n'
    probes = []
    for n in names:
        st = getattr(abi, n)
        src += f'printf("%zu\\n", sizeof({n}));\n'
        probes.append((n, None, C.sizeof(st)))
        for fname, _ in st._fields_:
            src += f'printf("%zu\\n", offsetof({n}, {fname}));\n'
            probes.append((n, fname, getattr(st, fname).offset))
    src += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "p"), os.path.join(d, "p.c")])
        out = subprocess.check_output([os.path.join(d, "p")], text=True).split()
    assert len(out) == len(probes)
    for (n, f, py), c in zip(probes, out):
        assert py == int(c), (n, f, py, c)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nptranscript_b200 import engine
    with pytest.raises(engine.NptError) as e:
        engine.Engine(engine.make_config())
    assert e.value.code == abi.NPT_ENODEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nptranscript_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert not re.search(r"liboracle|oracle\.cpp|oracle\.py|CDLL\([^)]*oracle|#include\s+[\"<][^\">]*oracle", txt), f
