"""CPU-side checks: the C-ABI library loads and exports every symbol include/lgs_b200.h declares."""
import os
import re

from my_lidar_graph_slam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lgs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lgs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in lgs_b200.h but not exported"


def test_binding_covers_header():
    assert sorted(capi.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_gpu():
    """Without a usable GPU the context must fail loudly (no silent CPU path)."""
    import ctypes as C
    h = C.c_void_p()
    rc = capi.lib().lgs_ctx_create(10_000, C.byref(h))   # no such device anywhere
    assert rc == 2 and not h.value


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary must be bindable from C (no C++ or torch types in the signatures)."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "lgs_b200.h"\nint main(void) { lgs_match_result r; r.found = 0; return r.found + LGS_OK; }\n')
    p = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), str(src)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
