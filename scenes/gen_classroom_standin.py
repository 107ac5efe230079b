"""Procedural stand-in for `scenes/classroom` (BASELINE.json config 3).

The reference ships only `scenes/classroom/scene.xml`; its 79 OBJ meshes and textures must be downloaded
(`scenes/classroom/how-to-obtain.txt`) and there is no network here. This script writes a LABELLED STAND-IN with the same
make-up -- a Mitsuba 0.5 XML scene (so it goes through the same tinyparser-mitsuba -> load_mitsuba_scene path), the
same integrator / sensor / sunsky blocks as the reference file, diffuse + roughplastic + (rough)conductor materials in
similar proportion, an interior lit by sun + sky through windows -- and geometry generated from a fixed recipe:
a 8 x 3.2 x 13 m room, 20 desks and chairs with tessellated tube legs, blackboard, shelves with books, ceiling lamps,
a globe. About 0.3 M triangles in ~40 shapes. Deterministic: no random numbers, only closed-form placement.

    python scenes/gen_classroom_standin.py [out_dir]      (default scenes/_generated/classroom_standin)
"""
import os
import sys

import numpy as np


class Mesh:
    """Indexed triangle mesh with per-vertex normal and uv."""

    def __init__(self):
        self.v, self.n, self.t, self.f = [], [], [], []
        self.count = 0

    def add(self, v, n, t, f):
        v, n, t, f = np.asarray(v, np.float64), np.asarray(n, np.float64), np.asarray(t, np.float64), np.asarray(f, np.int64)
        self.v.append(v), self.n.append(n), self.t.append(t), self.f.append(f + self.count)
        self.count += len(v)

    def tris(self):
        return sum(len(f) for f in self.f)

    def write(self, path, name):
        v, n, t, f = np.concatenate(self.v), np.concatenate(self.n), np.concatenate(self.t), np.concatenate(self.f) + 1
        with open(path, "w") as fh:
            fh.write(f"o {name}\n")
            np.savetxt(fh, v, fmt="v %.6f %.6f %.6f")
            np.savetxt(fh, t, fmt="vt %.6f %.6f")
            np.savetxt(fh, n, fmt="vn %.6f %.6f %.6f")
            idx = np.repeat(f, 3, axis=1)
            np.savetxt(fh, idx, fmt="f %d/%d/%d %d/%d/%d %d/%d/%d")


def grid_quad(mesh, origin, eu, ev, nu, nv):
    """Planar nu x nv grid spanning origin + s*eu + t*ev; normal = eu x ev."""
    origin, eu, ev = (np.asarray(a, np.float64) for a in (origin, eu, ev))
    s, t = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="ij")
    v = origin + s[..., None] * eu + t[..., None] * ev
    nrm = np.cross(eu, ev)
    nrm /= np.linalg.norm(nrm)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * (nv + 1) + j).ravel()
    b, c, d = a + (nv + 1), a + (nv + 1) + 1, a + 1
    f = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
    mesh.add(v.reshape(-1, 3), np.tile(nrm, ((nu + 1) * (nv + 1), 1)), np.stack([s, t], -1).reshape(-1, 2) * 4.0, f)


def box(mesh, lo, hi, sub=1):
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    grid_quad(mesh, lo, ey, ex, sub, sub)                 # -z
    grid_quad(mesh, lo + ez, ex, ey, sub, sub)            # +z
    grid_quad(mesh, lo, ez, ey, sub, sub)                 # -x
    grid_quad(mesh, lo + ex, ey, ez, sub, sub)            # +x
    grid_quad(mesh, lo, ex, ez, sub, sub)                 # -y
    grid_quad(mesh, lo + ey, ez, ex, sub, sub)            # +y


def tube(mesh, p0, p1, radius, seg=20, rings=6):
    p0, p1 = np.asarray(p0, np.float64), np.asarray(p1, np.float64)
    axis = p1 - p0
    length = np.linalg.norm(axis)
    w = axis / length
    u = np.cross(w, [1.0, 0, 0] if abs(w[0]) < 0.9 else [0, 1.0, 0])
    u /= np.linalg.norm(u)
    vv = np.cross(w, u)
    ang = np.linspace(0, 2 * np.pi, seg + 1)
    h = np.linspace(0, 1, rings + 1)
    A, H = np.meshgrid(ang, h, indexing="ij")
    nrm = np.cos(A)[..., None] * u + np.sin(A)[..., None] * vv
    v = p0 + H[..., None] * axis + radius * nrm
    i, j = np.meshgrid(np.arange(seg), np.arange(rings), indexing="ij")
    a = (i * (rings + 1) + j).ravel()
    b, c, d = a + (rings + 1), a + (rings + 1) + 1, a + 1
    f = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
    mesh.add(v.reshape(-1, 3), nrm.reshape(-1, 3), np.stack([A / (2 * np.pi), H], -1).reshape(-1, 2), f)


def sphere(mesh, center, radius, nu=64, nv=32):
    th = np.linspace(0, 2 * np.pi, nu + 1)
    ph = np.linspace(1e-3, np.pi - 1e-3, nv + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    nrm = np.stack([np.sin(P) * np.cos(T), np.cos(P), np.sin(P) * np.sin(T)], -1)
    v = np.asarray(center, np.float64) + radius * nrm
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * (nv + 1) + j).ravel()
    b, c, d = a + (nv + 1), a + (nv + 1) + 1, a + 1
    f = np.concatenate([np.stack([a, c, b], 1), np.stack([a, d, c], 1)])
    mesh.add(v.reshape(-1, 3), nrm.reshape(-1, 3), np.stack([T / (2 * np.pi), P / np.pi], -1).reshape(-1, 2), f)


def torus(mesh, center, R, r, nu=48, nv=20, axis=1):
    th = np.linspace(0, 2 * np.pi, nu + 1)
    ph = np.linspace(0, 2 * np.pi, nv + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    ring = np.stack([np.cos(T), np.zeros_like(T), np.sin(T)], -1)
    nrm = np.cos(P)[..., None] * ring + np.sin(P)[..., None] * np.array([0, 1.0, 0])
    v = R * ring + r * nrm
    if axis == 2:  # ring in the xy plane
        v, nrm = v[..., [0, 2, 1]], nrm[..., [0, 2, 1]]
    v = v + np.asarray(center, np.float64)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * (nv + 1) + j).ravel()
    b, c, d = a + (nv + 1), a + (nv + 1) + 1, a + 1
    f = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
    mesh.add(v.reshape(-1, 3), nrm.reshape(-1, 3), np.stack([T / (2 * np.pi), P / (2 * np.pi)], -1).reshape(-1, 2), f)


# (id, mitsuba type, colour, alpha or None)
MATERIALS = [
    ("Walls", "diffuse", (0.654071, 0.67408, 0.8), None), ("Floor", "diffuse", (0.55, 0.42, 0.28), None),
    ("Ceiling", "diffuse", (0.9, 0.9, 0.88), None), ("Blackboard", "diffuse", (0.03, 0.08, 0.05), None),
    ("Chalk", "diffuse", (0.9, 0.9, 0.9), None), ("BookRed", "diffuse", (0.8, 0.008214, 0.0), None),
    ("BookBlue", "diffuse", (0.05, 0.1, 0.6), None), ("BookGreen", "diffuse", (0.1, 0.5, 0.15), None),
    ("BookYellow", "diffuse", (0.779661, 0.653162, 0.349188), None), ("Shelf", "diffuse", (0.8, 0.644901, 0.412119), None),
    ("LampShade", "diffuse", (0.95, 0.95, 0.95), None), ("Poster", "diffuse", (0.7, 0.3, 0.3), None),
    ("Curtain", "diffuse", (0.85, 0.8, 0.7), None), ("Bin", "diffuse", (0.2, 0.2, 0.25), None),
    ("ClockFace", "diffuse", (1.0, 1.0, 1.0), None), ("Skirting", "diffuse", (0.3, 0.2, 0.12), None),
    ("DeskTop", "roughplastic", (0.647814, 0.5, 0.35), 0.1), ("ChairSeat", "roughplastic", (0.00631, 0.00631, 0.00631), 0.1),
    ("BoardFrame", "roughplastic", (1.0, 1.0, 1.0), 0.05), ("Globe", "roughplastic", (0.2, 0.4, 0.8), 0.05),
    ("TeacherDesk", "plastic", (0.4, 0.25, 0.12), None),
    ("DeskLegs", "roughconductor", (0.751534, 0.751534, 0.751534), 0.1), ("ChairLegs", "conductor", (1.0, 1.0, 1.0), None),
    ("WindowFrame", "roughconductor", (0.6, 0.6, 0.62), 0.2),
]


def build_meshes():
    m = {mid: Mesh() for mid, *_ in MATERIALS}
    X0, X1, Y0, Y1, Z0, Z1 = -4.5, 3.5, 0.0, 3.2, -6.0, 7.0
    # shell, inward-facing normals
    grid_quad(m["Floor"], (X0, Y0, Z0), (0, 0, Z1 - Z0), (X1 - X0, 0, 0), 96, 64)
    grid_quad(m["Ceiling"], (X0, Y1, Z0), (X1 - X0, 0, 0), (0, 0, Z1 - Z0), 48, 64)
    grid_quad(m["Walls"], (X0, Y0, Z0), (X1 - X0, 0, 0), (0, Y1 - Y0, 0), 64, 32)          # front wall (z = Z0)
    grid_quad(m["Walls"], (X1, Y0, Z1), (X0 - X1, 0, 0), (0, Y1 - Y0, 0), 64, 32)          # back wall
    grid_quad(m["Walls"], (X1, Y0, Z0), (0, 0, Z1 - Z0), (0, Y1 - Y0, 0), 96, 32)          # right wall
    # left wall with three window openings (z ranges), sill at 0.9 m, lintel at 2.6 m
    windows = [(-4.5, -1.5), (-0.5, 2.5), (3.5, 6.0)]
    edges = [Z0] + [z for w in windows for z in w] + [Z1]
    for k in range(0, len(edges), 2):  # solid piers
        z0, z1 = edges[k], edges[k + 1]
        grid_quad(m["Walls"], (X0, Y0, z1), (0, 0, z0 - z1), (0, Y1 - Y0, 0), max(2, int(8 * (z1 - z0))), 32)
    for z0, z1 in windows:
        grid_quad(m["Walls"], (X0, Y0, z1), (0, 0, z0 - z1), (0, 0.9, 0), 24, 9)          # below the sill
        grid_quad(m["Walls"], (X0, 2.6, z1), (0, 0, z0 - z1), (0, Y1 - 2.6, 0), 24, 6)     # above the lintel
        for zz in (z0, z1, 0.5 * (z0 + z1)):
            tube(m["WindowFrame"], (X0, 0.9, zz), (X0, 2.6, zz), 0.03, 16, 12)
        for yy in (0.9, 1.75, 2.6):
            tube(m["WindowFrame"], (X0, yy, z0), (X0, yy, z1), 0.03, 16, 16)
        box(m["Curtain"], (X0 + 0.05, 0.8, z1 - 0.35), (X0 + 0.12, 2.75, z1), 8)
    box(m["Skirting"], (X0, 0, Z0), (X1, 0.12, Z0 + 0.03), 4)
    box(m["Skirting"], (X1 - 0.03, 0, Z0), (X1, 0.12, Z1), 4)
    # blackboard on the front wall
    box(m["Blackboard"], (-3.0, 0.9, Z0 + 0.02), (2.0, 2.2, Z0 + 0.05), 16)
    for (a, b) in (((-3.05, 0.85), (2.05, 0.9)), ((-3.05, 2.2), (2.05, 2.25)), ((-3.05, 0.85), (-3.0, 2.25)), ((2.0, 0.85), (2.05, 2.25))):
        box(m["BoardFrame"], (a[0], a[1], Z0 + 0.02), (b[0], b[1], Z0 + 0.08), 4)
    for i in range(6):
        tube(m["Chalk"], (-2.5 + 0.25 * i, 0.92, Z0 + 0.1), (-2.42 + 0.25 * i, 0.92, Z0 + 0.1), 0.006, 12, 2)
    box(m["Poster"], (2.4, 1.2, Z0 + 0.02), (3.2, 2.3, Z0 + 0.03), 4)
    # desks and chairs: 4 columns x 5 rows
    for cx in range(4):
        for rz in range(5):
            x, z = -3.4 + 1.7 * cx, -3.2 + 1.75 * rz
            box(m["DeskTop"], (x, 0.72, z), (x + 1.1, 0.75, z + 0.6), 6)
            for dx, dz in ((0.05, 0.05), (1.05, 0.05), (0.05, 0.55), (1.05, 0.55)):
                tube(m["DeskLegs"], (x + dx, 0, z + dz), (x + dx, 0.72, z + dz), 0.018, 16, 8)
            tube(m["DeskLegs"], (x + 0.05, 0.15, z + 0.05), (x + 1.05, 0.15, z + 0.05), 0.012, 12, 8)
            cz = z + 0.85
            box(m["ChairSeat"], (x + 0.3, 0.44, cz), (x + 0.8, 0.47, cz + 0.45), 4)
            box(m["ChairSeat"], (x + 0.3, 0.7, cz + 0.42), (x + 0.8, 0.95, cz + 0.45), 4)
            for dx, dz in ((0.33, 0.03), (0.77, 0.03), (0.33, 0.42), (0.77, 0.42)):
                tube(m["ChairLegs"], (x + dx, 0, cz + dz), (x + dx, 0.95 if dz > 0.2 else 0.44, cz + dz), 0.012, 12, 8)
    # teacher's desk, globe, bin, clock
    box(m["TeacherDesk"], (-0.9, 0.0, -5.0), (0.9, 0.8, -4.2), 12)
    sphere(m["Globe"], (0.5, 1.05, -4.6), 0.22, 96, 48)
    torus(m["DeskLegs"], (0.5, 1.05, -4.6), 0.25, 0.008, 64, 10, axis=2)
    tube(m["DeskLegs"], (0.5, 0.8, -4.6), (0.5, 0.83, -4.6), 0.1, 24, 2)
    tube(m["Bin"], (2.9, 0, -5.4), (2.9, 0.4, -5.4), 0.16, 32, 8)
    tube(m["ClockFace"], (0.0, 2.75, Z0 + 0.02), (0.0, 2.75, Z0 + 0.05), 0.2, 48, 2)
    torus(m["WindowFrame"], (0.0, 2.75, Z0 + 0.05), 0.2, 0.015, 64, 12, axis=2)
    # shelves with books along the right wall
    for s in range(3):
        z = -4.5 + 3.6 * s
        box(m["Shelf"], (X1 - 0.4, 0, z), (X1 - 0.02, 1.8, z + 0.04), 4)
        box(m["Shelf"], (X1 - 0.4, 0, z + 2.4), (X1 - 0.02, 1.8, z + 2.44), 4)
        for lvl in range(5):
            y = 0.02 + 0.44 * lvl
            box(m["Shelf"], (X1 - 0.4, y, z), (X1 - 0.02, y + 0.03, z + 2.44), 4)
            for b in range(40):
                col = ("BookRed", "BookBlue", "BookGreen", "BookYellow")[(b * 7 + lvl * 3 + s) % 4]
                hgt = 0.24 + 0.03 * ((b * 5 + lvl) % 5)
                zz = z + 0.06 + 0.058 * b
                box(m[col], (X1 - 0.32 - 0.01 * (b % 3), y + 0.03, zz), (X1 - 0.06, y + 0.03 + hgt, zz + 0.05), 1)
    # ceiling lamps
    for lx in (-2.5, 1.5):
        for lz in (-3.5, 0.5, 4.5):
            box(m["LampShade"], (lx - 0.6, Y1 - 0.12, lz - 0.12), (lx + 0.6, Y1 - 0.04, lz + 0.12), 6)
            tube(m["LampShade"], (lx - 0.55, Y1 - 0.15, lz), (lx + 0.55, Y1 - 0.15, lz), 0.02, 16, 12)
    return {k: v for k, v in m.items() if v.count}


SCENE_HEAD = """<?xml version="1.0" encoding="utf-8"?>
<!-- PROCEDURAL STAND-IN for scenes/classroom (assets of the reference scene are not redistributable / not in the tree).
     Integrator, sensor and sunsky blocks are those of the reference's scenes/classroom/scene.xml. -->
<scene version="0.5.0" >
	<integrator type="path" >
		<integer name="maxDepth" value="17" />
		<boolean name="strictNormals" value="true" />
	</integrator>
	<sensor type="perspective" >
		<float name="fov" value="60" />
		<transform name="toWorld" >
			<matrix value="-0.988479 -0.00428443 0.151294 -1.69049 -9.42177e-010 -0.999599 -0.0283071 1.27158 -0.151355 0.027981 -0.988083 5.88653 0 0 0 1"/>
		</transform>
		<sampler type="sobol" >
			<integer name="sampleCount" value="64" />
		</sampler>
		<film type="ldrfilm" >
			<integer name="width" value="1280" />
			<integer name="height" value="720" />
		</film>
	</sensor>
"""
SCENE_TAIL = """	<emitter type="sunsky" >
		<vector name="sunDirection" x="-0.865804" y="0.916766" z="-0.276929" />
		<vector name="sunColor" x="0.98" y="0.82" z="0.30" />
		<vector name="skyColor" x="0.53" y="0.8" z="0.92" />
		<float name="sunScale" value="0.1" />
	</emitter>
</scene>
"""


def bsdf_xml(mid, kind, col, alpha):
    rgb = ", ".join(f"{c:g}" for c in col)
    if kind == "diffuse":
        inner = f'<bsdf type="diffuse" >\n\t\t\t<rgb name="reflectance" value="{rgb}"/>\n\t\t</bsdf>'
    elif kind in ("roughplastic", "plastic"):
        a = f'\n\t\t\t<float name="alpha" value="{alpha:g}" />' if alpha is not None else ""
        inner = (f'<bsdf type="{kind}" >{a}\n\t\t\t<float name="intIOR" value="1.5" />\n\t\t\t<float name="extIOR" value="1" />'
                 f'\n\t\t\t<rgb name="diffuseReflectance" value="{rgb}"/>\n\t\t</bsdf>')
    else:
        a = f'\n\t\t\t<float name="alpha" value="{alpha:g}" />' if alpha is not None else ""
        inner = (f'<bsdf type="{kind}" >{a}\n\t\t\t<float name="extEta" value="1" />\n\t\t\t<rgb name="specularReflectance" value="{rgb}"/>'
                 f'\n\t\t\t<rgb name="eta" value="1.65746, 0.880369, 0.521229"/>\n\t\t\t<rgb name="k" value="9.22387, 6.26952, 4.837"/>\n\t\t</bsdf>')
    return f'\t<bsdf type="twosided" id="{mid}" >\n\t\t{inner}\n\t</bsdf>\n'


def generate(out_dir):
    """Writes scene.xml + models/*.obj under out_dir (skips the work when already complete). Returns (xml path, #tris)."""
    xml_path = os.path.join(out_dir, "scene.xml")
    stamp = os.path.join(out_dir, ".complete")
    if os.path.exists(stamp) and os.path.exists(xml_path):
        return xml_path, int(open(stamp).read())
    os.makedirs(os.path.join(out_dir, "models"), exist_ok=True)
    meshes = build_meshes()
    xml = [SCENE_HEAD]
    for mid, kind, col, alpha in MATERIALS:
        if mid in meshes:
            xml.append(bsdf_xml(mid, kind, col, alpha))
    total = 0
    for k, (mid, mesh) in enumerate(meshes.items()):
        fn = f"models/Mesh{k:03d}.obj"
        mesh.write(os.path.join(out_dir, fn), mid)
        total += mesh.tris()
        xml.append(f'\t<shape type="obj" >\n\t\t<string name="filename" value="{fn}" />\n\t\t<transform name="toWorld" >\n'
                   f'\t\t\t<matrix value="1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1"/>\n\t\t</transform>\n\t\t<ref id="{mid}" />\n\t</shape>\n')
    xml.append(SCENE_TAIL)
    with open(xml_path, "w") as fh:
        fh.write("".join(xml))
    with open(stamp, "w") as fh:
        fh.write(str(total))
    return xml_path, total


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "_generated", "classroom_standin")
    path, n = generate(out)
    print(path, n, "triangles")
