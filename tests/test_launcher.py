"""The Boost-free launcher (adapters/launcher): config C1 from a CARMEN log + launcher_settings_default.json
through the reference's own CarmenLogReader, LidarGraphSlam, front end and back end.

CPU part: the unmodified default settings (reference classes only) run here without a GPU.
GPU part: the same log with the B200 type strings must write the same pose graph, bit for bit."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "adapters", "_build", "lgs_slam_launch")
SETTINGS = os.path.join(ROOT, "tests", "golden", "launcher_settings_default.json")
needs_exe = pytest.mark.skipif(not os.path.exists(EXE), reason="adapters/_build/lgs_slam_launch not built "
                               "(needs the reference tree at build time)")


def _log(tmp_path, n=160, seed=3):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_carmen_log
    path = str(tmp_path / "synthetic.log")
    make_carmen_log.write_log(path, n, seed)
    return path


def _run(log, out, *sets):
    cmd = [EXE, log, SETTINGS, out, "--set", "Backend.PoseGraphOptimizerType=None"]
    for kv in sets:
        cmd += ["--set", kv]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    return p


@needs_exe
def test_default_settings_run_from_a_carmen_log(tmp_path):
    log = _log(tmp_path)
    p = _run(log, str(tmp_path / "cpu"))
    assert p.returncode == 0, p.stderr[-2000:]
    info = json.loads(p.stdout.strip().splitlines()[-1])
    assert info["scans"] == 160 and info["frames"] >= 20 and info["nodes"] == info["frames"]
    assert info["scan_matcher"] == "RealTimeCorrelative" and info["loop_detector"] == "BranchBound"
    poses = open(tmp_path / "cpu.poses.txt").read().splitlines()
    assert len(poses) == info["nodes"] and poses[0].split()[1:] == ["0", "0", "0"]
    # the estimated trajectory follows the drive (0.1 m per scan) and not just the drifting odometry
    x, y = (float(v) for v in poses[-1].split()[1:3])
    assert 1.0 < (x * x + y * y) ** 0.5 < 20.0
    q = _run(log, str(tmp_path / "cpu2"))
    assert q.returncode == 0 and open(tmp_path / "cpu2.poses.txt").read() == "\n".join(poses) + "\n"


@needs_exe
def test_unavailable_types_fail_loudly(tmp_path):
    log = _log(tmp_path, n=12)
    p = subprocess.run([EXE, log, SETTINGS, str(tmp_path / "x")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode != 0 and "sparse solvers" in p.stderr            # "LM" without Eigen
    p = _run(log, str(tmp_path / "y"), "Frontend.LocalSlam.ScanMatcherType=NoSuchMatcher")
    assert p.returncode != 0 and "NoSuchMatcher" in p.stderr
    p = _run(log, str(tmp_path / "z"), "Backend.LoopDetectorType=Nope")
    assert p.returncode != 0 and "Nope" in p.stderr


@needs_exe
@pytest.mark.gpu
def test_cuda_type_strings_write_the_same_pose_graph(tmp_path):
    log = _log(tmp_path, n=420, seed=5)
    a = _run(log, str(tmp_path / "cpu"))
    b = _run(log, str(tmp_path / "gpu"), "Frontend.LocalSlam.ScanMatcherType=RealTimeCorrelativeCuda",
             "Backend.LoopDetectorType=BranchBoundCuda")
    assert a.returncode == 0, a.stderr[-2000:]
    assert b.returncode == 0, b.stderr[-2000:]
    ia, ib = (json.loads(p.stdout.strip().splitlines()[-1]) for p in (a, b))
    assert ib["scan_matcher"] == "RealTimeCorrelativeCuda" and ib["loop_detector"] == "BranchBoundCuda"
    assert ia["frames"] == ib["frames"] >= 60
    assert open(tmp_path / "cpu.poses.txt").read() == open(tmp_path / "gpu.poses.txt").read()
    assert open(tmp_path / "cpu.edges.txt").read() == open(tmp_path / "gpu.edges.txt").read()
