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


def test_grid_search_offsets_replay_the_reference_loops():
    """lgs_gs_offsets (host arithmetic only): the accumulating loops of scan_matcher_grid_search.cpp:74-76,
    whose lengths depend on the rounding of the running sum -- checked against the loop lengths the
    unmodified reference produced for tests/golden/scene_gs.npz."""
    import numpy as np
    gold = np.load(os.path.join(ROOT, "tests", "golden", "scene_gs.npz"))
    params = [dict(range_x=0.6, range_y=0.5, range_theta=0.12, step_x=0.05, step_y=0.05, step_theta=0.01),
              dict(range_x=0.4, range_y=0.4, range_theta=0.1, step_x=0.03, step_y=0.07, step_theta=0.013),
              dict(range_x=0.3, range_y=0.3, range_theta=0.05, step_x=0.1, step_y=0.1, step_theta=0.005)]
    for n, ints in enumerate(gold["ints"]):
        p = params[n % 3]
        got = [len(capi.gs_offsets(p["range_" + a], p["step_" + a])) for a in ("x", "y", "theta")]
        assert got == list(ints[4:7])
    for rng_, step in ((2.0, 0.05), (0.5, 0.005), (0.12, 0.01), (1.0, 0.3), (0.0, 0.1)):
        want, d = [], -rng_ / 2.0
        while d <= rng_ / 2.0:
            want.append(d)
            d += step
        assert capi.gs_offsets(rng_, step).tolist() == want
    assert len(capi.gs_offsets(0.12, 0.01)) == 12          # not 13: the running sum overshoots 0.06
