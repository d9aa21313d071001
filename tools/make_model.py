#!/usr/bin/env python
"""Derive the mesh-dependent compile-time constants of template/mushr.em.xml from the reference's STL
assets (build container only; the GPU box has no reference tree).

    python tools/make_model.py [--from-golden tests/golden/mujoco_golden.npz]

With --from-golden the chassis CoM / inertia and the softener sphere are taken from constants captured from a real
MuJoCo compile (tests/golden/make_mujoco_golden.py) instead of the restated rule.
MuJoCo computes these when it compiles the MJCF; `mujoco` is absent here, so the
"legacy" mesh-inertia rule (the default of the mesh `inertia` attribute in the pinned 3.2.x / 3.3.x
releases) is restated: pyramids from the area-weighted face centroid to every triangle, |volume| each.

Outputs (committed):
  oracle/mushr_mesh.h                      constants for the CPU oracle
  ft_grandprix_b200/csrc/mushr_mesh.h      the same numbers for the CUDA product
  tests/golden/mushr_mesh.json             the same numbers for tests
"""
import json, os, struct, sys
import numpy as np

REF = os.environ.get("FTGP_REFERENCE", "/root/reference")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def load_stl(p):
    b = open(p, "rb").read()
    n = struct.unpack("<I", b[80:84])[0]
    if 84 + 50 * n == len(b):
        a = np.frombuffer(b[84:84 + 50 * n], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
        return a["v"].astype(np.float64)
    import re
    v = np.array([[float(x) for x in m] for m in re.findall(rb"vertex\s+(\S+)\s+(\S+)\s+(\S+)", b)])
    return v.reshape(-1, 3, 3)


def legacy_mesh(tri):
    """-> volume, CoM (mesh frame), inertia tensor about CoM at unit density (mesh axes)."""
    D, E, F = tri[:, 0], tri[:, 1], tri[:, 2]
    nrm = np.cross(E - D, F - D)
    area = 0.5 * np.linalg.norm(nrm, axis=1)
    nrm = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
    cen = (D + E + F) / 3
    facecen = (area[:, None] * cen).sum(0) / area.sum()
    vol = np.abs(((cen - facecen) * nrm).sum(1) * area / 3)
    com = (vol[:, None] * (cen * 0.75 + facecen * 0.25)).sum(0) / vol.sum()
    D, E, F = D - com, E - com, F - com
    cen = (D + E + F) / 3
    vol = np.abs((cen * nrm).sum(1) * area / 3)
    P = np.zeros((3, 3))
    for a in range(3):
        for b in range(3):
            P[a, b] = (vol / 20 * (2 * (D[:, a] * D[:, b] + E[:, a] * E[:, b] + F[:, a] * F[:, b])
                                   + D[:, a] * E[:, b] + D[:, b] * E[:, a] + D[:, a] * F[:, b] + D[:, b] * F[:, a]
                                   + E[:, a] * F[:, b] + E[:, b] * F[:, a])).sum()
    I = np.eye(3) * np.trace(P) - P
    return vol.sum(), com, I


def hull_vertices(pts):
    from scipy.spatial import ConvexHull
    u = np.unique(np.round(pts, 9), axis=0)
    return u[ConvexHull(u).vertices]


def quat2mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def from_golden(path):
    """Chassis CoM / inertia and the softener sphere from constants captured from a real MuJoCo compile
    (tests/golden/make_mujoco_golden.py: const_body_*, const_geom_*).  Body 1 = `car #0` = chassis mesh + lidar
    cylinder; the cylinder is analytic (density 1000, r 0.03, half height 0.015 at (-0.0525, 0, 0.0575),
    mushr.em.xml:108), so the mesh's share is what is left."""
    g = np.load(path, allow_pickle=False)
    mb, ipos = float(g["const_body_mass"][1]), g["const_body_ipos"][1]
    R = quat2mat(g["const_body_iquat"][1])
    Ib = R @ np.diag(g["const_body_inertia"][1]) @ R.T                    # about ipos, body axes
    lr, lh = 0.030, 0.015
    lm = 1000.0 * np.pi * lr * lr * 2 * lh
    lc = np.array([-0.0525, 0.0, 0.065 - lh / 2])
    Il = np.diag([lm * (3 * lr * lr + 4 * lh * lh) / 12] * 2 + [lm * lr * lr / 2])
    mass = mb - lm
    com = (mb * ipos - lm * lc) / mass
    shift = lambda m, d: m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    I = Ib - Il - shift(lm, lc - ipos) - shift(mass, com - ipos)          # chassis about its own CoM
    sid = int(g["const_softener_geom_ids"][0])
    return mass, com, I, float(g["const_geom_size"][sid][0]), g["const_geom_pos"][sid]


def main():
    golden = None
    if "--from-golden" in sys.argv:
        golden = from_golden(sys.argv[sys.argv.index("--from-golden") + 1])
    base = load_stl(os.path.join(REF, "template/meshes/simple_base_nano.stl")) * 0.5          # mushr.em.xml:38
    wheel = load_stl(os.path.join(REF, "template/meshes/mushr_wheel.stl")) * (0.5 * 1.3)       # mushr.em.xml:39
    # chassis geom: mass 3.542137 at geom pos (0, 0, 0.5 * 0.094655) (mushr.em.xml:119)
    vol, com, I = legacy_mesh(base)
    mass = 3.542137
    I = I * (mass / vol)
    gpos = np.array([0, 0, 0.5 * 0.094655])
    chassis = {"mass": mass, "com": (com + gpos).tolist(), "inertia": I.tolist(), "volume": vol,
               "hull": (hull_vertices(base.reshape(-1, 3)) + gpos).tolist()}
    # softener: sphere fitted to the wheel mesh's equivalent inertia box, x fitscale 2 (mushr.em.xml:66)
    wvol, wcom, wI = legacy_mesh(wheel)
    ev = np.sort(np.linalg.eigvalsh(wI))[::-1]
    # equivalent box half sizes from principal moments (mass = volume at unit density)
    ev = np.linalg.eigvalsh(wI)
    hs = [0.5 * np.sqrt(6 * (ev[(k + 1) % 3] + ev[(k + 2) % 3] - ev[k]) / wvol) for k in range(3)]
    radius = float(np.mean(hs)) * 2.0
    rule = "MuJoCo legacy mesh inertia restated (parity unpinned: mujoco absent)"
    if golden is not None:                                   # constants of a real MuJoCo compile win over the restated rule
        gm, gcom, gI, gr, gc = golden
        print("restated vs MuJoCo: chassis com", chassis["com"], gcom.tolist(), "inertia diag", np.diag(I), np.diag(gI),
              "softener radius", radius, gr, file=sys.stderr)
        assert abs(gm - mass) < 1e-6, (gm, mass)
        I, radius, wcom = gI, gr, np.asarray(gc)
        chassis["com"], chassis["inertia"] = gcom.tolist(), gI.tolist()
        rule = "captured from a MuJoCo compile (tests/golden/mujoco_golden.npz)"
    soft = {"radius": radius, "center": np.asarray(wcom).tolist(), "mass": 1e-5, "box_half": [float(h) for h in hs]}
    out = {"chassis": chassis, "softener": soft, "rule": rule}
    json.dump(out, open(os.path.join(ROOT, "tests/golden/mushr_mesh.json"), "w"), indent=1)
    H = chassis["hull"]
    body = ["/* GENERATED by tools/make_model.py from template/meshes/ STL files -- do not edit.",
            f" * Mesh-derived constants of template/mushr.em.xml: {rule}. */",
            f"#define MUSHR_CHASSIS_MASS {mass!r}",
            "#define MUSHR_CHASSIS_COM {%s}" % ", ".join(repr(float(x)) for x in chassis["com"]),
            "#define MUSHR_CHASSIS_INERTIA {%s}" % ", ".join(repr(float(x)) for x in I.reshape(-1)),
            f"#define MUSHR_SOFTENER_RADIUS {radius!r}",
            "#define MUSHR_SOFTENER_CENTER {%s}" % ", ".join(repr(float(x)) for x in wcom),
            f"#define MUSHR_CHASSIS_NTRI {len(base)}",
            "/* the chassis mesh's own triangles in the car frame (STL x 0.5 + geom pos): what mj_ray tests (SURVEY B.10) */",
            "#define MUSHR_CHASSIS_TRI {%s}" % ", ".join("{%s}" % ", ".join(repr(float(x)) for x in (t + gpos).reshape(-1)) for t in base),
            f"#define MUSHR_CHASSIS_NHULL {len(H)}",
            "#define MUSHR_CHASSIS_HULL {%s}" % ", ".join("{%r, %r, %r}" % tuple(float(x) for x in v) for v in H), ""]
    for p in ("oracle/mushr_mesh.h", "ft_grandprix_b200/csrc/mushr_mesh.h"):
        open(os.path.join(ROOT, p), "w").write("\n".join(body))
    print(json.dumps({k: out[k] for k in ("softener",)}, indent=1))
    print("chassis com", chassis["com"], "\ninertia\n", I, "\nvolume", vol, "nhull", len(H))


if __name__ == "__main__":
    main()
