"""The C-ABI library loads on a CPU-only box and exports every symbol include/tslam.h declares."""
import ctypes as C
import os
import re

import pytest

from taichislam_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "tslam.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tslam_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_loads():
    from taichislam_b200 import build
    build.build()
    L = capi.load()
    assert L.tslam_abi_version() == capi.ABI_VERSION


def test_every_declared_symbol_is_exported_and_bound():
    L = capi.load()
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"libtslam.so lacks {n}"
        assert n in capi.SIGNATURES, f"_capi.SIGNATURES lacks {n}"
    assert sorted(capi.SIGNATURES) == names


def test_struct_layout_matches_header():
    # field order/count of the ctypes mirrors vs the header text
    txt = open(os.path.join(ROOT, "include", "tslam.h")).read()
    body = re.search(r"typedef struct tslam_tsdf_config \{(.*?)\} tslam_tsdf_config_t;", txt, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ty, names = decl.split(None, 1)
        fields += [n.strip() for n in names.split(",")]
    assert fields == [f[0] for f in capi.TsdfConfig._fields_]


def test_no_cpu_fallback_without_gpu():
    L = capi.load()
    if L.tslam_device_count() > 0:
        pytest.skip("GPU present")
    cfg = capi.TsdfConfig(0.05, 64, 64, 10, 0.3, 10, 2, 1, 1, 0, 0, 1, -0.3, 1.8, 0, 0, 0, 0)
    h = C.c_void_p()
    assert L.tslam_tsdf_create(C.byref(cfg), C.byref(h)) == capi.E_NOGPU
    with pytest.raises(capi.TslamError):
        capi.require_gpu()


def test_product_never_imports_oracle():
    """The product path must not import, link, load or call anything under oracle/."""
    pkg = os.path.join(ROOT, "taichislam_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|libtslam_oracle|\borc_[a-z]+|oracle[/\\.]oracle|tslam_oracle", re.M)
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not bad.search(src), f"{f} reaches into oracle/"
