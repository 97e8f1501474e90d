"""The C-ABI boundary: struct layouts match the headers byte for byte, and both shared
libraries load and export every symbol the headers declare (no compute, no GPU)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def declared(header):
    src = open(os.path.join(INCLUDE, header)).read()
    return re.findall(r"RTB_API\s+[\w\s\*]+?\b(rtbh?_\w+)\s*\(", src)


def test_struct_sizes_match_the_headers(rtb, tmp_path):
    names = list(rtb.abi.STRUCT_SIZES)
    prog = '#include <stdio.h>\n#include "rtb_host.h"\nint main(void){\n' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "return 0;}\n"
    src = tmp_path / "sizes.c"
    src.write_text(prog)
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", INCLUDE, "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = dict(line.split() for line in out.strip().splitlines())
    for n in names:
        assert int(got[n]) == rtb.abi.STRUCT_SIZES[n], n
    assert rtb.abi.STRUCT_SIZES["rtb_view"] == 88            # View.cs:8-14: 7 float3 + float


def test_headers_compile_as_c(tmp_path):
    src = tmp_path / "c.c"
    src.write_text('#include "rtb.h"\n#include "rtb_host.h"\nint main(void){return RTB_ABI_VERSION == 2 ? 0 : 1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", INCLUDE, "-o", str(tmp_path / "c"), str(src)], check=True)


def test_plugin_exports_every_declared_symbol(rtb):
    names = declared("rtb.h")
    assert set(names) == set(rtb.plugin.EXPORTS) and len(names) >= 15
    path = rtb.build.plugin_lib_path()
    if not os.path.exists(path):
        rtb.build.build_plugin()
    L = C.CDLL(path)
    for n in names:
        assert hasattr(L, n), n
    assert rtb.plugin.lib().rtb_abi_version() == rtb.abi.ABI_VERSION == 2


def test_host_lib_exports_every_declared_symbol(rtb):
    names = declared("rtb_host.h")
    assert len(names) >= 9
    L = C.CDLL(rtb.build.build_host())
    for n in names:
        assert hasattr(L, n), n


def test_plugin_has_no_cpu_fallback(rtb):
    """Without a CUDA device rtb_create must fail loudly (and with one, this test is moot)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rtb.plugin.RtbError) as e:
        rtb.plugin.Context(0)
    assert e.value.code >= rtb.abi.RTB_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "raytracing-in-one-weekend_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle_lib" not in text and "oracle/" not in text.replace("CPU oracle", ""), f
