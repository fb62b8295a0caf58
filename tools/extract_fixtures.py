#!/usr/bin/env python3
"""Regenerate the committed fixtures under tests/golden/ from the read-only reference checkout.

Runs only where /root/reference exists (the build container).  Nothing at test/bench time reads the
reference; the tests read the files this script wrote.

Outputs
  tests/golden/scene_constants.json   masses / radii / box sizes parsed from the URDFs, the racket hull
                                      outline (2-D convex hull of racket.stl in the link y-z plane), the
                                      AABB-box inertia Bullet recomputes for the racket compound
  tests/golden/ppo_swing_policy.npz   weights of backup_models/ppo_swing.zip (SB3 1.8.0 MlpPolicy)
  tests/golden/ppo_swing_monitor.json the 100-episode Monitor buffer + `_last_obs` + PPO hyper-parameters
                                      stored in the same zip (the only numbers we hold that real PyBullet produced)
  tests/golden/es_swing_weights.npz   backup_models/es_swing.dat as a plain float32 vector

Reference sources read (all under /root/reference):
  tennisbot/resources/{racket,ball,court,simplegoal}.urdf, racket.stl, backup_models/*.
"""
import base64
import io
import json
import pickle
import struct
import sys
import warnings
import xml.etree.ElementTree as ET
import zipfile
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
RES = REF / "tennisbot" / "resources"

URDF_MARGIN = 0.001  # PyBullet's default collision margin for URDF shapes [R]


def _floats(s):
    return [float(x) for x in s.split()]


def hull2d(points):
    """Andrew monotone chain, CCW, no collinear points kept."""
    pts = sorted(set(map(tuple, points)))

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return lower[:-1] + upper[:-1]


def racket_constants():
    raw = (RES / "racket.stl").read_bytes()
    ntri = struct.unpack("<I", raw[80:84])[0]
    assert len(raw) == 84 + 50 * ntri
    rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
    tris = np.frombuffer(raw[84:], dtype=rec, count=ntri)
    verts = np.unique(tris["v"].reshape(-1, 3).astype(np.float64), axis=0)
    xs = np.unique(verts[:, 0])
    assert len(xs) == 2, "racket.stl is expected to be a plate extruded along x"
    poly = hull2d(verts[:, 1:3].tolist())
    # rotate so the list starts at the bottom-right handle corner (max y among the min-z vertices)
    zmin = min(p[1] for p in poly)
    start = max((i for i, p in enumerate(poly) if abs(p[1] - zmin) < 1e-6), key=lambda i: poly[i][0])
    poly = poly[start:] + poly[:start]
    area = 0.5 * sum(poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[(i + 1) % len(poly)][0] * poly[i][1]
                     for i in range(len(poly)))
    assert area > 0

    link = ET.parse(RES / "racket.urdf").getroot().find("link")
    inertial = link.find("inertial")
    mass = float(inertial.find("mass").get("value"))
    com = _floats(inertial.find("origin").get("xyz"))
    lo, hi = verts.min(0) - URDF_MARGIN, verts.max(0) + URDF_MARGIN
    ext = hi - lo
    inertia = [mass / 12.0 * (ext[1] ** 2 + ext[2] ** 2),
               mass / 12.0 * (ext[0] ** 2 + ext[2] ** 2),
               mass / 12.0 * (ext[0] ** 2 + ext[1] ** 2)]
    return {
        "mass": mass,
        "com_in_link": com,
        "half_thickness_x": float(xs[1]),
        "x_planes": xs.tolist(),
        "outline_yz_link": poly,           # CCW, link frame (subtract com_in_link[2] from z for the COM frame)
        "outline_area": area,
        "stl_triangles": int(ntri),
        "stl_unique_vertices": int(len(verts)),
        "aabb_extent_with_margin": ext.tolist(),
        "inertia_aabb_box": inertia,       # what btCompoundShape::calculateLocalInertia yields [R]
        "urdf_inertia_ignored": [float(inertial.find("inertia").get(k)) for k in ("ixx", "iyy", "izz")],
    }


def scene_constants():
    ball = ET.parse(RES / "ball.urdf").getroot().find("link")
    court = ET.parse(RES / "court.urdf").getroot().find("link")
    goal = ET.parse(RES / "simplegoal.urdf").getroot().find("link")
    boxes = [_floats(c.find("geometry").find("box").get("size")) for c in court.findall("collision")]
    cyl = goal.find("collision").find("geometry").find("cylinder")
    r_ball = float(ball.find("collision").find("geometry").find("sphere").get("radius"))
    m_ball = float(ball.find("inertial").find("mass").get("value"))
    return {
        "source": "tools/extract_fixtures.py over /root/reference/tennisbot/resources",
        "urdf_margin": URDF_MARGIN,
        "ball": {"radius": r_ball, "mass": m_ball, "inertia_sphere": 0.4 * m_ball * r_ball ** 2},
        "court": {"floor_box_size": boxes[0], "net_box_size": boxes[1]},
        "goal": {"radius": float(cyl.get("radius")), "length": float(cyl.get("length")), "prism_sides": 32},
        "racket": racket_constants(),
    }


def ppo_fixture():
    import torch

    z = zipfile.ZipFile(REF / "backup_models" / "ppo_swing.zip")
    sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True, map_location="cpu")
    np.savez(OUT / "ppo_swing_policy.npz", **{k.replace(".", "__"): v.numpy() for k, v in sd.items()})
    data = json.loads(z.read("data"))

    def unpickle(key):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return pickle.loads(base64.b64decode(data[key][":serialized:"]))

    eps = unpickle("ep_info_buffer")
    keep = ("n_envs", "n_steps", "batch_size", "n_epochs", "gamma", "gae_lambda", "ent_coef", "vf_coef",
            "max_grad_norm", "learning_rate", "num_timesteps", "_n_updates", "normalize_advantage")
    mon = {
        "source": "backup_models/ppo_swing.zip (SB3 %s)" % z.read("_stable_baselines3_version").decode().strip(),
        "system_info": z.read("system_info.txt").decode(),
        "episode_returns": [float(e["r"]) for e in eps],
        "episode_lengths": [int(e["l"]) for e in eps],
        "last_obs": np.asarray(unpickle("_last_obs"), dtype=np.float64).reshape(-1).tolist(),
        "hyper": {k: data[k] for k in keep},
    }
    (OUT / "ppo_swing_monitor.json").write_text(json.dumps(mon, indent=1))


def es_fixture():
    import torch

    w = torch.load(REF / "backup_models" / "es_swing.dat", weights_only=False)
    np.savez(OUT / "es_swing_weights.npz", weights=np.asarray(w, dtype=np.float32))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "scene_constants.json").write_text(json.dumps(scene_constants(), indent=1))
    ppo_fixture()
    es_fixture()
    print("fixtures written to", OUT)


if __name__ == "__main__":
    main()
