"""Write a synthetic CARMEN log in the FLASER format the reference's CarmenLogReader parses
(io/carmen/carmen_reader.cpp:319-395: `FLASER n r_0 .. r_{n-1} lx ly ltheta rx ry rtheta ts host logts`):
180 beams, 1 degree apart from -90 degrees (GuessAngleIncrement(180) = pi / 180, :484-490), an Intel-lab-like
drive through axis-aligned rooms with drifting odometry (config C1).

  python tools/make_carmen_log.py out.log [n_scans=400] [seed=1]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from my_lidar_graph_slam_b200 import synth  # noqa: E402


def write_log(path, n_scans=400, seed=1, step=0.1):
    world = synth.RoomsWorld(30.0, 5.0, seed=seed)
    angles = -np.pi / 2 + np.arange(180) * (np.pi / 180.0)
    traj = synth.trajectory(world, n_scans, step=step, seed=seed)
    rng = np.random.default_rng(seed + 1)
    first = traj[0].copy()
    with open(path, "w") as f:
        f.write("# synthetic CARMEN log (tools/make_carmen_log.py): FLASER, 180 beams\n")
        for k, p in enumerate(traj):
            ranges = np.minimum(synth.make_scan(world, p, angles, rng), 79.0)
            # the log's poses are ODOMETRY: the true motion relative to the first pose plus a slow drift
            d = p - first
            c, s = np.cos(-first[2]), np.sin(-first[2])
            odom = np.array([c * d[0] - s * d[1] + 0.0015 * k, s * d[0] + c * d[1] - 0.001 * k, d[2] + 0.0004 * k])
            t = 0.2 * k
            f.write("FLASER 180 " + " ".join(f"{r:.4f}" for r in ranges))
            f.write(" {0!r} {1!r} {2!r} {0!r} {1!r} {2!r} {3:.3f} synth {3:.3f}\n".format(
                float(odom[0]), float(odom[1]), float(odom[2]), t))
    return traj


if __name__ == "__main__":
    out = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    write_log(out, n, seed)
    print(f"wrote {n} scans to {out}")
