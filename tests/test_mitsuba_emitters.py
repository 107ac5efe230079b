"""Mitsuba `rectangle` shapes and nested `area` emitters. The reference's loader drops them
(LumenScene.cpp:538-540, MitsubaParser.cpp:121-142) and so does lumen_b200 by default; with LUMEN_B200_MITSUBA_AREA_EMITTERS=1
they become two-triangle meshes / emissive materials, i.e. area lights of Lumen's own kind."""
import os

import numpy as np
import pytest

from helpers import bits_equal, pixel_agreement
from lumen_b200 import host

FLOOR_OBJ = """o floor
v -4 0 -4
v 4 0 -4
v 4 0 4
v -4 0 4
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 1 0
f 1/1/1 3/3/1 2/2/1
f 1/1/1 4/4/1 3/3/1
"""

SCENE_XML = """<?xml version="1.0" encoding="utf-8"?>
<scene version="0.5.0" >
	<integrator type="path" >
		<integer name="maxDepth" value="5" />
	</integrator>
	<sensor type="perspective" >
		<float name="fov" value="70" />
		<transform name="toWorld" >
			<matrix value="1 0 0 0 0 0.8 0.6 3 0 -0.6 0.8 5 0 0 0 1"/>
		</transform>
	</sensor>
	<bsdf type="twosided" id="grey" >
		<bsdf type="diffuse" >
			<rgb name="reflectance" value="0.6, 0.6, 0.6"/>
		</bsdf>
	</bsdf>
	<shape type="obj" >
		<string name="filename" value="floor.obj" />
		<transform name="toWorld" >
			<matrix value="1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1"/>
		</transform>
		<ref id="grey" />
	</shape>
	<shape type="rectangle" >
		<transform name="toWorld" >
			<matrix value="0.7 0 0 0.5 0 0 0.7 2.5 0 -0.7 0 -0.25 0 0 0 1"/>
		</transform>
		<ref id="grey" />
		<emitter type="area" >
			<rgb name="radiance" value="12, 10, 8"/>
		</emitter>
	</shape>
</scene>
"""


@pytest.fixture()
def xml_scene(tmp_path):
    (tmp_path / "floor.obj").write_text(FLOOR_OBJ)
    (tmp_path / "scene.xml").write_text(SCENE_XML)
    return str(tmp_path / "scene.xml")


def test_default_drops_rectangles_like_the_reference(xml_scene, monkeypatch):
    monkeypatch.delenv("LUMEN_B200_MITSUBA_AREA_EMITTERS", raising=False)
    sc = host.Scene(xml_scene, 64, 48)
    assert (sc.info.n_prim_meshes, sc.info.n_triangles, sc.info.n_lights, sc.info.total_light_triangle_cnt) == (1, 2, 0, 0)


def test_opt_in_makes_an_area_light(xml_scene, monkeypatch):
    from oracle import pyoracle as po
    monkeypatch.setenv("LUMEN_B200_MITSUBA_AREA_EMITTERS", "1")
    sc = host.Scene(xml_scene, 64, 48)
    assert (sc.info.n_prim_meshes, sc.info.n_triangles, sc.info.n_lights, sc.info.total_light_triangle_cnt) == (2, 4, 1, 2)
    assert sc.info.n_materials == 2  # the referenced bsdf and its emissive clone
    # 1.4 x 1.4 rectangle: the loader's area sum (LumenScene.cpp:115-133) sees the baked world-space triangles
    assert abs(sc.info.total_light_area - 1.96) < 1e-4
    img, st = po.OracleScene(sc).render(sc.make_pc(5, True), sc.make_ubo(), 0, 4)
    assert st.rays_probe > 0  # MIS probes exist only for area lights
    assert img[..., :3].mean() > 0.05 and np.isfinite(img).all()


@pytest.mark.gpu
def test_opt_in_scene_gpu_equals_oracle(xml_scene, monkeypatch, device):
    from oracle import pyoracle as po
    monkeypatch.setenv("LUMEN_B200_MITSUBA_AREA_EMITTERS", "1")
    sc = host.Scene(xml_scene, 160, 120)
    orc = po.OracleScene(sc)
    pc, ubo = sc.make_pc(5, True), sc.make_ubo()
    device.upload_scene(sc.desc)
    device.build_accel()
    device.init(160, 120, 3)
    device.render(pc, ubo, 0, 8)
    gpu, gs = device.download(), device.stats()
    cpu, cs = orc.render(pc, ubo, 0, 8)
    assert (gs.rays_closest, gs.rays_shadow, gs.rays_probe) == (cs.rays_closest, cs.rays_shadow, cs.rays_probe)
    assert pixel_agreement(gpu, cpu) >= 0.999 and bits_equal(gpu, cpu).mean() >= 0.999


@pytest.mark.gpu
def test_scene_without_any_light_is_sky_only_and_matches(xml_scene, monkeypatch, device):
    """Default loading drops the rectangle emitter, which leaves a scene with NO light: the reference still takes a light sample
    (of a zeroed Light: num_lights = 0, light_triangle_count = 0 -> a contribution of 0 / inf = 0) and a shadow ray per bounce.
    Both sides do exactly that; only the constant sky lights the floor."""
    from oracle import pyoracle as po
    monkeypatch.delenv("LUMEN_B200_MITSUBA_AREA_EMITTERS", raising=False)
    sc = host.Scene(xml_scene, 96, 72)
    assert sc.info.n_lights == 0
    pc, ubo = sc.make_pc(5, True), sc.make_ubo()
    pc.sky_col[0], pc.sky_col[1], pc.sky_col[2] = 0.5, 0.6, 0.7
    device.upload_scene(sc.desc)
    device.build_accel()
    device.init(96, 72, 2)
    device.render(pc, ubo, 0, 4)
    gpu, gs = device.download(), device.stats()
    cpu, cs = po.OracleScene(sc).render(pc, ubo, 0, 4)
    assert (gs.rays_closest, gs.rays_shadow, gs.rays_probe) == (cs.rays_closest, cs.rays_shadow, cs.rays_probe)
    assert gs.rays_shadow > 0 and gpu[..., :3].mean() > 0.05
    assert bits_equal(gpu, cpu).mean() >= 0.999
