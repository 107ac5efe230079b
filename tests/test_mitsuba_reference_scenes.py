"""the Mitsuba loader on the reference's OWN inputs, scenes/classroom/scene.xml (config 3) and
scenes/bedroom/scene.xml (config 4). The 79 / 69 OBJ meshes and the textures those files name are not distributed with the
reference (how-to-obtain.txt), so the test writes a one-triangle placeholder for every `models/*.obj` and a 2 x 2 image for every
texture next to a copy of the XML; everything else -- bsdf unwrapping and the bsdf -> Material mapping, shape -> bsdf references,
toWorld matrices, camera, sunsky, max depth -- comes from the real file and is checked against an INDEPENDENT reading of the XML
(ElementTree) mapped by the rules of LumenScene.cpp:514-690 and MitsubaParser.cpp:5-147."""
import ctypes as C
import math
import os
import shutil
import struct
import xml.etree.ElementTree as ET
import zlib

import numpy as np
import pytest

from conftest import ROOT
from lumen_b200 import host
from lumen_b200._ctypes_types import Light, Material, PrimMeshInfo

DIFFUSE, GLASS, CONDUCTOR, PRINCIPLED = 1, 4, 16, 32
F_DIFFUSE, F_SPECULAR, F_GLOSSY, F_REFLECTION, F_TRANSMISSION = 1, 2, 4, 8, 16


def write_png(path, w=2, h=2, rgb=(200, 120, 40)):
    raw = b"".join(b"\x00" + bytes(rgb) * w for _ in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def stage(name, tmp_path):
    """copy scenes/<name>/scene.xml and create placeholders for every file it references; returns (xml path, parsed tree)"""
    src = os.path.join(ROOT, "scenes", name, "scene.xml")
    dst = tmp_path / name
    (dst / "models").mkdir(parents=True)
    (dst / "textures").mkdir()
    shutil.copy(src, dst / "scene.xml")
    tree = ET.parse(src)
    for s in tree.getroot().iter("string"):
        if s.get("name") != "filename":
            continue
        f = dst / s.get("value")
        if f.suffix == ".obj":
            k = int("".join(ch for ch in f.stem if ch.isdigit()) or 0)
            f.write_text(f"o {f.stem}\nv 0 0 {k}\nv 1 0 {k}\nv 0 1 {k}\nvn 0 0 1\nvt 0 0\nvt 1 0\nvt 0 1\nf 1/1/1 2/2/1 3/3/1\n")
        else:
            write_png(str(f))  # stb_image goes by content, not by extension
    return str(dst / "scene.xml"), tree


def expected_materials(tree):
    """LumenScene.cpp:607-672 applied to the XML's top-level <bsdf> elements, in document order."""
    out = []
    for b in tree.getroot().findall("bsdf"):
        name = b.get("id")
        while b.get("type") in ("twosided", "mask") and b.find("bsdf") is not None:  # MitsubaParser.cpp:43-46
            b = b.find("bsdf")
        t = b.get("type")
        albedo, rough, ior, texture = (1.0, 1.0, 1.0), 0.0, 1.0, None
        for c in b:
            n = c.get("name")
            # MitsubaParser.cpp:50-57; tinyparser turns camelCase names into snake_case, so diffuseReflectance / specularReflectance
            # both contain "reflectance" (no bsdf of these files carries two of them; eta / k colours are skipped)
            if c.tag == "rgb" and "reflectance" in n.lower():
                albedo = tuple(float(x) for x in c.get("value").split(","))
            if c.tag == "float" and n == "alpha":
                rough = math.sqrt(np.float32(float(c.get("value"))))  # :59-61
            if c.tag == "float" and n == "intIOR":
                ior = float(c.get("value"))
            if c.tag == "texture":
                texture = [s.get("value") for s in c.iter("string") if s.get("name") == "filename"][0]
        m = dict(name=name, type=t, albedo=albedo, roughness=rough, ior=ior, texture=texture)
        if t == "diffuse":
            m.update(bsdf_type=DIFFUSE, props=F_DIFFUSE | F_REFLECTION)
        elif t in ("roughplastic", "plastic", "roughdielectric", "dielectric"):
            props = (F_DIFFUSE | F_REFLECTION if rough < 1.0 else 0) | (F_TRANSMISSION if ior != 1.0 else 0) | (F_GLOSSY if np.float32(rough) > np.float32(0.08) else F_SPECULAR)
            m.update(bsdf_type=PRINCIPLED, props=props)
            if t in ("roughdielectric", "dielectric"):
                m.update(spec_trans=1.0, metallic=0.0, subsurface=0.0, thin=0)
            else:
                m.update(spec_trans=0.5, metallic=1.0, subsurface=0.1, thin=1)
        elif t in ("conductor", "roughconductor"):
            m.update(bsdf_type=CONDUCTOR, props=F_REFLECTION | (F_GLOSSY if np.float32(rough) > np.float32(0.08) else F_SPECULAR))
        else:
            raise AssertionError(f"bsdf type {t} not expected in these files")
        out.append(m)
    return out


def materials_of(sc):
    return (Material * sc.info.n_materials).from_address(sc.desc.materials)


def matrix_of(elem):
    v = [float(x) for x in elem.find("transform").find("matrix").get("value").split()]
    return np.float32(v).reshape(4, 4)  # row-major in the file


@pytest.mark.parametrize("name,n_shapes,n_obj,counts,sun_scale", [
    ("classroom", 79, 79, {"diffuse": 36, "plastic": 5, "conductor": 3, "dielectric": 0}, 0.1),
    ("bedroom", 72, 70, {"diffuse": 20, "plastic": 2, "conductor": 6, "dielectric": 3}, 0.5),
])
def test_reference_scene_xml_tables(name, n_shapes, n_obj, counts, sun_scale, tmp_path):
    path, tree = stage(name, tmp_path)
    sc = host.Scene(path, 192, 108)
    root = tree.getroot()
    shapes = root.findall("shape")
    assert len(shapes) == n_shapes and sum(1 for s in shapes if s.get("type") == "obj") == n_obj

    # ---- materials (LumenScene.cpp:607-672): one per top-level bsdf, in file order
    want = expected_materials(tree)
    kinds = {"diffuse": 0, "plastic": 0, "conductor": 0, "dielectric": 0}
    for m in want:
        kinds["plastic" if "plastic" in m["type"] else "dielectric" if "dielectric" in m["type"] else "conductor" if "conductor" in m["type"] else "diffuse"] += 1
    assert kinds == counts  # classroom = 36 diffuse, 5 plastic -> principled, 3 conductor
    mats = materials_of(sc)
    assert sc.info.n_materials == len(want)
    n_tex = 0
    for got, m in zip(mats, want):
        assert got.bsdf_type == m["bsdf_type"], m["name"]
        assert got.bsdf_props == m["props"], m["name"]
        assert got.roughness == np.float32(m["roughness"]), m["name"]
        if m["texture"]:
            assert got.texture_id == n_tex  # textures are numbered in material order (:609-613)
            n_tex += 1
        else:
            assert got.texture_id == -1
        if m["bsdf_type"] == PRINCIPLED:
            assert (got.metallic, got.spec_trans, got.thin) == (np.float32(m["metallic"]), np.float32(m["spec_trans"]), m["thin"]), m["name"]
            assert got.subsurface == np.float32(m["subsurface"]) and got.ior == np.float32(m["ior"])
            # make_default_principled (:597-606) zeroes the remaining Disney parameters
            assert (got.specular_tint, got.sheen_tint, got.clearcoat, got.clearcoat_gloss, got.sheen) == (0, 0, 0, 0, 0)
        if m["bsdf_type"] == CONDUCTOR:
            # reflectance_to_conductor_eta_k (LumenScene.cpp:47-50): eta = 1, k = 2 sqrt(R) / sqrt(max(1 - R, 0.001)); the Mitsuba
            # path does not clamp R (the JSON path does, :418), so specularReflectance 1 gives k = 2 / sqrt(0.001)
            R = np.float32(m["albedo"])
            k = np.float32(2.0) * np.sqrt(R) / np.sqrt(np.maximum(np.float32(1.0) - R, np.float32(0.001)))
            assert np.allclose(list(got.albedo), 1.0) and np.allclose(list(got.k), k, rtol=1e-6)
        elif m["bsdf_type"] != CONDUCTOR:
            assert np.allclose(list(got.albedo), m["albedo"], rtol=1e-6), m["name"]
    assert sc.info.n_textures == n_tex
    assert sc.info.bsdf_types == (DIFFUSE | PRINCIPLED | CONDUCTOR)

    # ---- shapes: file-less shapes (bedroom's two rectangle emitters) are dropped (LumenScene.cpp:538-540, Q11: the reference keeps
    # zero-sized trailing prim-mesh slots for them; they hold no triangle); every obj shape keeps its bsdf reference and toWorld matrix
    obj_shapes = [s for s in shapes if s.get("type") == "obj"]
    assert sc.info.n_prim_meshes == len(obj_shapes) and sc.info.n_triangles == len(obj_shapes)
    infos = (PrimMeshInfo * sc.info.n_prim_meshes).from_address(sc.desc.prim_infos)
    world = np.ctypeslib.as_array(C.cast(sc.desc.world_matrices, C.POINTER(C.c_float)), shape=(sc.info.n_prim_meshes, 16))
    ids = [m["name"] for m in want]
    for i, s in enumerate(obj_shapes):
        assert infos[i].material_index == ids.index(s.find("ref").get("id")), i
        assert infos[i].index_offset == 3 * i and infos[i].vertex_offset == 3 * i
        M = matrix_of(s)  # column-major storage of the same matrix (MitsubaParser.cpp:98-106 transposes the row-major source)
        assert (world[i].reshape(4, 4).T == M).all(), i

    # ---- sunsky -> one directional, delta light, L = 100 * sun_color * sun_scale (MitsubaParser.cpp:121-142, LumenScene.cpp:679-689)
    sun = [e for e in root.findall("emitter") if e.get("type") == "sunsky"][0]
    vec = {v.get("name"): np.float32([float(v.get(a)) for a in "xyz"]) for v in sun.findall("vector")}
    assert float(sun.find("float").get("value")) == sun_scale
    lights = (Light * sc.info.n_lights).from_address(sc.desc.lights)
    assert sc.info.n_lights == 1 and sc.info.dir_light_idx == 0 and sc.info.total_light_triangle_cnt == 1
    L = lights[0]
    assert L.light_flags == (3 | (1 << 5))
    assert np.allclose(list(L.L), np.float32(100.0) * (vec["sunColor"] * np.float32(sun_scale)), rtol=1e-6)
    assert (np.float32(list(L.pos)) == vec["sunDirection"]).all() and list(L.to) == [0, 0, 0]
    assert np.allclose(list(sc.info.sky_col), vec["skyColor"])
    assert L.world_radius == sc.info.world_radius and L.world_radius > 0

    # ---- integrator + camera: maxDepth from the file; fov / 2 (LumenScene.cpp:532); position = last column of toWorld (Camera.h:103)
    assert sc.info.integrator == b"path"
    assert sc.info.path_length == int(root.find("integrator").find("integer").get("value"))
    sensor = root.find("sensor")
    fov = float(sensor.find("float").get("value")) / 2
    ubo = sc.make_ubo()
    proj = np.float32(list(ubo.projection)).reshape(4, 4)  # column-major: proj[c][r]
    f = 1.0 / math.tan(math.radians(fov) / 2)
    assert np.isclose(proj[1][1], -f, rtol=1e-5) and np.isclose(proj[0][0], f / (192 / 108), rtol=1e-5)  # Camera.h:109-123 (Y flip)
    inv_view = np.float32(list(ubo.inv_view)).reshape(4, 4)
    assert np.allclose(inv_view[3][:3], matrix_of(sensor)[:3, 3], atol=1e-5)
    pc = sc.make_pc(8, True)
    assert (pc.num_lights, pc.light_triangle_count, pc.dir_light_idx, pc.max_depth) == (1, 1, 0, 8)


def test_bedroom_rectangle_emitters_are_an_opt_in(tmp_path, monkeypatch):
    """The reference drops bedroom's two lamp rectangles (file == "" -> continue) and never parses <emitter type="area">; with
    LUMEN_B200_MITSUBA_AREA_EMITTERS=1 they become two-triangle area lights with the XML's radiance."""
    path, tree = stage("bedroom", tmp_path)
    base = host.Scene(path, 64, 36)
    monkeypatch.setenv("LUMEN_B200_MITSUBA_AREA_EMITTERS", "1")
    sc = host.Scene(path, 64, 36)
    assert sc.info.n_prim_meshes == base.info.n_prim_meshes + 2 and sc.info.n_triangles == base.info.n_triangles + 4
    assert sc.info.n_lights == 3 and sc.info.total_light_triangle_cnt == 2 + 2 + 1 and sc.info.n_materials == base.info.n_materials + 2
    mats = materials_of(sc)
    for m in list(mats)[-2:]:
        assert np.allclose(list(m.emissive_factor), 16.4648) and m.bsdf_type == DIFFUSE and list(m.albedo) == [0, 0, 0]
