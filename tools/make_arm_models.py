"""Generate csrc/arm_models_data.h -- the geometric model used by the arm collision kernel AND its oracle.

Run in the build container (needs /root/reference for the URDFs and STL meshes):
    python tools/make_arm_models.py

The reference decides arm collisions with PyBullet contact generation on convex hulls of the URDF meshes
(environment/kuka_env.py:354-370); PyBullet is not installable here and unpinned in the reference, so parity
with it is UNPINNED.  This script defines OUR model instead, from the reference's own
assets: the joint chain (origins, axes, limits) is read from the URDF, and the convex hull of every
link mesh (what PyBullet collides with) is filled with K inscribed spheres (greedy set cover of the hull interior).  A state
is "free" iff every sphere keeps more than MARGIN from every obstacle box (and, for two arms, from every
sphere of the other arm).  Only data is emitted (no reference code is copied).
"""
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = os.environ.get("GNNMP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gnn_motion_planning_b200", "csrc",
                   "arm_models_data.h")
K_SPHERES = 12
MARGIN = 0.0      # inscribed spheres already under-approximate the hulls; calibrated on the reference problem sets


def rpy_matrix(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def read_stl(path):
    data = open(path, "rb").read()
    n = struct.unpack("<I", data[80:84])[0]
    if 84 + 50 * n == len(data):
        rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
        return rec["v"].reshape(-1, 3).astype(np.float64)
    verts = []
    for line in data.decode(errors="ignore").splitlines():
        t = line.split()
        if len(t) == 4 and t[0] == "vertex":
            verts.append([float(x) for x in t[1:]])
    return np.array(verts)


def fit_spheres(pts, k, n_samples=6000, seed=0):
    """K spheres INSCRIBED in the convex hull of the mesh (PyBullet loads URDF meshes as convex hulls): candidate
    centres are uniform samples of the hull interior, a candidate's radius is its distance to the nearest hull
    facet, and spheres are picked greedily to cover the most not-yet-covered interior samples (set cover).
    The union of the spheres is therefore a subset of the hull: the model never reports a collision PyBullet's
    hull would not, and misses only shallow contacts at uncovered corners."""
    from scipy.spatial import ConvexHull
    pts = np.unique(np.round(pts, 6), axis=0)
    hull = ConvexHull(pts)
    A, b = hull.equations[:, :3], hull.equations[:, 3]
    rng = np.random.default_rng(seed)
    lo, hi = pts.min(0), pts.max(0)
    samples = np.zeros((0, 3))
    while len(samples) < n_samples:
        cand = rng.uniform(lo, hi, (4 * n_samples, 3))
        inside = np.all(cand @ A.T + b <= 0, axis=1)
        samples = np.concatenate([samples, cand[inside]])
    samples = samples[:n_samples]
    rad = np.min(-(samples @ A.T + b), axis=1)            # inscribed radius at every sample
    covered = np.zeros(len(samples), bool)
    centers, radii = [], []
    d2 = None
    for _ in range(k):
        best, best_gain = -1, -1
        # evaluate the candidates with the largest radii (cheap, good enough for a greedy cover)
        order = np.argsort(-rad)[:600]
        for i in order:
            gain = np.count_nonzero(~covered & (((samples - samples[i]) ** 2).sum(1) <= rad[i] ** 2))
            if gain > best_gain:
                best, best_gain = i, gain
        if best_gain <= 0:
            break
        centers.append(samples[best].copy())
        radii.append(rad[best])
        covered |= ((samples - samples[best]) ** 2).sum(1) <= rad[best] ** 2
        rad = np.where(covered, -1.0, rad) if False else rad
    fit_spheres.last_coverage = covered.mean()
    return np.array(centers), np.array(radii)


def chain_from_urdf(urdf, mesh_dir):
    root = ET.parse(urdf).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = [j for j in root.findall("joint")]
    assert all(j.get("type") == "revolute" for j in joints), "only all-revolute serial chains are handled"
    order, parent = [], joints[0].find("parent").get("link")
    base = parent
    rem = list(joints)
    while rem:
        nxt = [j for j in rem if j.find("parent").get("link") == parent]
        assert len(nxt) == 1, "not a serial chain"
        order.append(nxt[0])
        rem.remove(nxt[0])
        parent = nxt[0].find("child").get("link")
    J = []
    for j in order:
        o = j.find("origin")
        xyz = [float(x) for x in o.get("xyz").split()]
        rpy = [float(x) for x in o.get("rpy").split()]
        axis = [float(x) for x in j.find("axis").get("xyz").split()]
        lim = j.find("limit")
        J.append(dict(R=rpy_matrix(*rpy), t=np.array(xyz), axis=np.array(axis), lo=float(lim.get("lower")), hi=float(lim.get("upper"))))
    spheres = []  # (frame index: 0 = base link, j+1 = child of joint j)
    names = [base] + [j.find("child").get("link") for j in order]
    for fi, name in enumerate(names):
        col = links[name].find("collision")
        if col is None:
            continue
        mesh = col.find("geometry").find("mesh")
        if mesh is None:
            continue
        path = os.path.join(os.path.dirname(urdf), mesh.get("filename"))
        if not os.path.exists(path):
            print("  (no mesh file for %s: link has no collision geometry, as in PyBullet)" % name)
            continue
        o = col.find("origin")
        Rm = rpy_matrix(*[float(x) for x in o.get("rpy").split()])
        tm = np.array([float(x) for x in o.get("xyz").split()])
        pts = read_stl(path) @ Rm.T + tm
        c, r = fit_spheres(pts, K_SPHERES)
        for ci, ri in zip(c, r):
            spheres.append((fi, ci, ri))
        print("  %-18s %6d verts -> %d inscribed spheres, radii %.3f..%.3f, hull volume covered %.0f%%" % (
            name, len(pts), len(c), r.min(), r.max(), 100 * fit_spheres.last_coverage))
    return J, spheres


def tree_chain_from_urdf(urdf, skip_links=()):
    """General URDF tree with fixed joints (ur5.urdf): the revolute joints along the path root -> tip become the chain;
    fixed joints are folded into the next revolute joint's origin (or, after the last one, into the sphere centres).
    Frame 0 = everything rigidly attached to the root, frame k = child side of the k-th revolute joint."""
    root = ET.parse(urdf).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    children = {}
    for j in joints:
        children.setdefault(j.find("parent").get("link"), []).append(j)
    child_names = {j.find("child").get("link") for j in joints}
    base = [n for n in links if n not in child_names][0]

    def origin_of(el):
        o = el.find("origin")
        xyz = [float(x) for x in (o.get("xyz") if o is not None and o.get("xyz") else "0 0 0").split()]
        rpy = [float(x) for x in (o.get("rpy") if o is not None and o.get("rpy") else "0 0 0").split()]
        return rpy_matrix(*rpy), np.array(xyz)

    J, spheres = [], []

    def visit(link, frame, R, t):      # (R, t): pose of `link` in its movable frame
        col = links[link].find("collision")
        if col is not None and link not in skip_links:
            mesh = col.find("geometry").find("mesh")
            if mesh is not None:
                Rm, tm = origin_of(col)
                pts = read_stl(os.path.join(os.path.dirname(urdf), mesh.get("filename"))) @ Rm.T + tm
                pts = pts @ R.T + t
                c, r = fit_spheres(pts, K_SPHERES)
                for ci, ri in zip(c, r):
                    spheres.append((frame, ci, ri))
                print("  %-18s frame %d %6d verts -> %d inscribed spheres, radii %.3f..%.3f, hull volume covered %.0f%%" % (
                    link, frame, len(pts), len(c), r.min(), r.max(), 100 * fit_spheres.last_coverage))
        for j in children.get(link, []):
            Rj, tj = origin_of(j)
            Rc, tc = R @ Rj, R @ tj + t
            child = j.find("child").get("link")
            if j.get("type") == "revolute":
                lim = j.find("limit")
                axis = np.array([float(x) for x in j.find("axis").get("xyz").split()])
                J.append(dict(R=Rc, t=tc, axis=axis, lo=float(lim.get("lower")), hi=float(lim.get("upper"))))
                assert len(J) == frame + 1, "branching kinematic trees are not handled"
                visit(child, frame + 1, np.eye(3), np.zeros(3))
            else:
                visit(child, frame, Rc, tc)

    visit(base, 0, np.eye(3), np.zeros(3))
    spheres.sort(key=lambda s: s[0])
    return J, spheres


def snake_model():
    """environment/snake.urdf + SnakeEnv.set_config (snake_env.py:118-135): a planar free base (x, y at height 0.5, yaw) and
    four revolute z joints; links are spheres (r 0.08) and capsules (length 0.44, r 0.05, axis = local y).  Capsules are
    filled with inscribed spheres along their axis.  Reference quirk kept: yaw AND joint 3 both read config[3], config[6]
    is unused (qidx below)."""
    root = ET.parse(os.path.join(REF, "environment/snake.urdf")).getroot()
    radii = {l.get("name"): l.find("collision").find("geometry") for l in root.findall("link")}
    ball_r = float(radii["ball_0"].find("sphere").get("radius"))
    cap = radii["link_1"].find("capsule")
    cap_len, cap_r = float(cap.get("length")), float(cap.get("radius"))
    seg = float(root.find("joint").find("origin").get("xyz").split()[1])          # -0.4: every joint steps -0.4 in y
    eye = np.eye(3)
    J = [dict(R=eye, t=np.array([0, 0, 0.5]), axis=np.array([1.0, 0, 0]), lo=-9.0, hi=9.0, type=1, qidx=0),   # base x (SnakeEnv.height)
         dict(R=eye, t=np.zeros(3), axis=np.array([0, 1.0, 0]), lo=-9.0, hi=9.0, type=1, qidx=1),              # base y
         dict(R=eye, t=np.zeros(3), axis=np.array([0, 0, 1.0]), lo=-np.pi, hi=np.pi, type=0, qidx=3)]          # base yaw <- config[3]
    spheres = []

    def add_block(frame):        # ball at the frame origin, capsule centred one step further along -y
        spheres.append((frame, np.zeros(3), ball_r))
        n = 9
        for y in np.linspace(-cap_len / 2, cap_len / 2, n):
            spheres.append((frame, np.array([0.0, seg + y, 0.0]), cap_r))

    add_block(3)
    for k, qi in enumerate((2, 3, 4, 5)):      # joints 1,3,5,7 <- config[2..5]   (snake_env.py:127-128)
        J.append(dict(R=eye, t=np.array([0.0, 2 * seg, 0.0]), axis=np.array([0, 0, 1.0]), lo=-np.pi, hi=np.pi, type=0, qidx=qi))
        add_block(4 + k)
    return J, spheres


def emit_model(f, name, J, spheres):
    f.write("static const GmpJoint %s_joints[] = {\n" % name)
    for idx, j in enumerate(J):
        Rt = ", ".join("%.17g" % x for x in j["R"].reshape(-1))
        f.write("  {{%s}, {%.17g, %.17g, %.17g}, {%.17g, %.17g, %.17g}, %.17g, %.17g, %d, %d},\n" % (
            Rt, *j["t"], *j["axis"], j["lo"], j["hi"], j.get("type", 0), j.get("qidx", idx)))
    f.write("};\nstatic const GmpSphere %s_spheres[] = {\n" % name)
    for fi, c, r in spheres:
        f.write("  {%d, {%.17g, %.17g, %.17g}, %.17g},\n" % (fi, *c, r))
    f.write("};\n\n")


def main():
    models = {}
    print("kuka7 <- kuka_iiwa/model_0.urdf")
    models["kuka7"] = chain_from_urdf(os.path.join(REF, "kuka_iiwa/model_0.urdf"), os.path.join(REF, "kuka_iiwa"))
    print("kuka13 <- kuka_iiwa/model_3.urdf")
    models["kuka13"] = chain_from_urdf(os.path.join(REF, "kuka_iiwa/model_3.urdf"), os.path.join(REF, "kuka_iiwa"))
    print("ur5 <- ur5/ur5.urdf")
    models["ur5"] = tree_chain_from_urdf(os.path.join(REF, "ur5/ur5.urdf"), skip_links=("ee_link",))
    print("snake7 <- environment/snake.urdf")
    models["snake7"] = snake_model()
    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/make_arm_models.py from the reference's URDF + STL assets -- data only.\n"
                "// Shared by csrc/arm.cu (the kernel) and oracle/arm.c (its checker): the geometric SPEC of the arm model.\n"
                "// Joint chain: frame 0 = base link; frame j+1 = frame j * [R|t]_j * Rot(axis_j, q_j).\n"
                "#pragma once\n\n"
                "// type: 0 = revolute about `axis`, 1 = prismatic along `axis`; qidx = which state component drives the joint\n"
                "typedef struct { double R[9]; double t[3]; double axis[3]; double lo, hi; int type, qidx; } GmpJoint;\n"
                "typedef struct { int frame; double c[3]; double r; } GmpSphere;\n"
                "#define GMP_ARM_MARGIN %.17g\n#define GMP_ARM_MAX_JOINTS 14\n#define GMP_ARM_MAX_SPHERES 176\n\n" % MARGIN)
        for name, (J, S) in models.items():
            emit_model(f, "gmp_" + name, J, S)
        f.write("// model ids: 0 = kuka7 (KukaEnv, kuka_env.py), 1 = kuka14 (Kuka2Env: two kuka7 chains based at x = -0.5 / +0.5,\n"
                "// kuka_2arm_env.py:58-59, arm-arm contacts included), 2 = kuka13 (KukaEnv with model_3.urdf),\n"
                "// 3 = ur5 (UR5Env, ur5_env.py:104-127: self collision between links that are not directly connected, ground plane z = 0\n"
                "// except for the base link; the 1 cm ee_link box is dropped)\n"
                "// 4 = snake7 (SnakeEnv, snake_env.py: planar base + 4 joints, self collision incl. directly connected links; state limits\n"
                "// (-9,9)^2 x (-pi,pi)^5 from snake_env.py:55; the ground plane never touches a robot riding at z = 0.5)\n"
                "#define GMP_ARM_KUKA7 0\n#define GMP_ARM_KUKA14 1\n#define GMP_ARM_KUKA13 2\n#define GMP_ARM_UR5 3\n#define GMP_ARM_SNAKE7 4\n"
                "#define GMP_ARM_NUM_MODELS 5\n"
                "static const double gmp_snake7_lo[7] = {-9, -9, %.17g, %.17g, %.17g, %.17g, %.17g};\n"
                "static const double gmp_snake7_hi[7] = {9, 9, %.17g, %.17g, %.17g, %.17g, %.17g};\n" % ((-np.pi,) * 5 + (np.pi,) * 5))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
