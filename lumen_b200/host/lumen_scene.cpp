// Scene ingest: behaviour-for-behaviour equivalent of Lumen's LumenScene without Vulkan. See lumen_scene.h.
// Third-party parsers are the ones Lumen itself vendors (tinyobj, nlohmann json, tinyparser-mitsuba, stb_image).
#include "lumen_scene.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>

#define GLM_ENABLE_EXPERIMENTAL
#include <glm/gtc/matrix_transform.hpp>
#include <glm/gtc/quaternion.hpp>
#include <glm/gtc/type_ptr.hpp>
#include <glm/gtx/euler_angles.hpp>
#include <glm/gtx/matrix_decompose.hpp>
#include <glm/gtx/quaternion.hpp>

#include <tinygltf/json.hpp>
#define TINYOBJLOADER_IMPLEMENTATION
#include <tiny_obj_loader.h>
#define STB_IMAGE_IMPLEMENTATION
#include <stb_image/stb_image.h>
#include <mitsuba_parser/tinyparser-mitsuba.h>

namespace lmh {
using json = nlohmann::json;

namespace {

bool ends_with(const std::string& s, const std::string& e) { return e.size() <= s.size() && std::equal(e.rbegin(), e.rend(), s.rbegin()); }

float jf(json& j, const char* key, float def) { return j[key].is_null() ? def : (float)j[key]; }
uint32_t ju(json& j, const char* key, uint32_t def) { return j[key].is_null() ? def : (uint32_t)j[key]; }
glm::vec3 jv3(const json& a) { return glm::vec3((float)a[0], (float)a[1], (float)a[2]); }
void put3(float* dst, const glm::vec3& v) { dst[0] = v.x, dst[1] = v.y, dst[2] = v.z; }
glm::vec3 get3(const float* p) { return glm::vec3(p[0], p[1], p[2]); }

// LumenScene.cpp:47-50
void reflectance_to_conductor_eta_k(const glm::vec3& reflectance, glm::vec3& eta, glm::vec3& k) {
	eta = glm::vec3(1.0f);
	k = 2.0f * glm::sqrt(reflectance) / glm::sqrt(glm::max(glm::vec3(1.0f) - reflectance, 0.001f));
}

// BBox.h:2-65
struct Bbox {
	glm::vec3 mn{FLT_MAX}, mx{-FLT_MAX};
	void insert(const glm::vec3& v) { mn = glm::min(mn, v), mx = glm::max(mx, v); }
	bool is_empty() const { return mn == glm::vec3(FLT_MAX) || mx == glm::vec3(-FLT_MAX); }
	bool is_volume() const { return (mn.x < mx.x) + (mn.y < mx.y) + (mn.z < mx.z) == 3; }
};

std::string lower(std::string s) {
	std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
	return s;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ camera
void PerspectiveCamera::make_projection() {
	projection = glm::mat4(1.0f);
	projection[0][0] = 1 / (aspect_ratio * tanf(glm::radians(fov / 2)));
	projection[1][1] = -1 / (tanf(glm::radians(fov / 2)));
	projection[2][2] = cam_far / (cam_near - cam_far);
	projection[2][3] = -1;
	projection[3][2] = cam_near * cam_far / (cam_near - cam_far);
}

static glm::vec3 euler_from_view(const glm::mat4& view) {
	glm::vec3 scale, translation, skew;
	glm::quat q;
	glm::vec4 perspective;
	glm::decompose(view, scale, q, translation, skew, perspective);
	glm::vec3 rot{};
	glm::extractEulerAngleXYZ(glm::toMat4(q), rot.x, rot.y, rot.z);
	rot *= 180. / glm::pi<float>();
	return rot;
}

void PerspectiveCamera::init_lookat(float fov_deg, float aspect, const glm::vec3& dir, const glm::vec3& pos) {
	fov = fov_deg, aspect_ratio = aspect;
	make_projection();
	position = pos;
	view = glm::lookAtLH(position, position + dir, glm::vec3(0, 1, 0));
	rotation = euler_from_view(view);
}

void PerspectiveCamera::init_matrix(float fov_deg, const glm::mat4& cam_matrix, float aspect) {
	fov = fov_deg, aspect_ratio = aspect;
	make_projection();
	view = glm::inverse(cam_matrix);
	rotation = euler_from_view(view);
	position = glm::vec3(cam_matrix[0][3], cam_matrix[1][3], cam_matrix[2][3]);
}

void PerspectiveCamera::update_view_matrix() {
	glm::mat4 res{1};
	res = glm::translate(res, position);
	res = glm::rotate(res, glm::radians(rotation.y), glm::vec3(0, 1, 0));
	res = glm::rotate(res, glm::radians(rotation.x), glm::vec3(1, 0, 0));
	res = glm::rotate(res, glm::radians(rotation.z), glm::vec3(0, 0, 1));
	view = glm::inverse(res);
}

// ------------------------------------------------------------------------------------------------ OBJ ingest
// LumenScene.cpp:287-326 / 563-593: de-index every face corner into its own vertex.
void Scene::append_obj_shape(const void* attrib_p, const void* shape_p, PrimMesh& pm) {
	const auto& attrib = *static_cast<const tinyobj::attrib_t*>(attrib_p);
	const auto& shape = *static_cast<const tinyobj::shape_t*>(shape_p);
	pm.first_idx = (uint32_t)indices.size();
	pm.vtx_offset = (uint32_t)positions.size();
	pm.name = shape.name;
	pm.idx_count = (uint32_t)shape.mesh.indices.size();
	pm.vtx_count = (uint32_t)shape.mesh.num_face_vertices.size();
	glm::vec3 min_vtx(FLT_MAX), max_vtx(-FLT_MAX);
	uint32_t index_offset = 0, idx_val = 0;
	for (uint32_t f = 0; f < shape.mesh.num_face_vertices.size(); f++) {
		for (uint32_t v = 0; v < 3; v++) {
			const tinyobj::index_t idx = shape.mesh.indices[index_offset + v];
			indices.push_back(idx_val++);
			const glm::vec3 p(attrib.vertices[3 * (size_t)idx.vertex_index + 0], attrib.vertices[3 * (size_t)idx.vertex_index + 1],
							  attrib.vertices[3 * (size_t)idx.vertex_index + 2]);
			positions.push_back(p);
			min_vtx = glm::min(p, min_vtx);
			max_vtx = glm::max(p, max_vtx);
			// The reference appends normals / texcoords only when present and then indexes them per position
			// (LumenScene.cpp:312-322 vs :178-183), i.e. it requires vn and vt on every face corner.
			if (idx.normal_index < 0 || idx.texcoord_index < 0)
				throw std::runtime_error("OBJ shape '" + shape.name + "' lacks vn/vt on a face corner; Lumen's loader requires both");
			normals.emplace_back(attrib.normals[3 * (size_t)idx.normal_index + 0], attrib.normals[3 * (size_t)idx.normal_index + 1],
								 attrib.normals[3 * (size_t)idx.normal_index + 2]);
			texcoords0.emplace_back(attrib.texcoords[2 * (size_t)idx.texcoord_index + 0], attrib.texcoords[2 * (size_t)idx.texcoord_index + 1]);
		}
		index_offset += 3;
	}
	pm.min_pos = min_vtx;
	pm.max_pos = max_vtx;
}

// ------------------------------------------------------------------------------------------------ Lumen JSON
void Scene::load_lumen_scene(const std::string& path) {
	const std::string root = path.substr(0, path.find_last_of("/\\") + 1);
	std::ifstream in(path);
	if (!in) throw std::runtime_error("cannot open scene file " + path);
	json j;
	in >> j;

	auto& integrator = j["integrator"];
	config.integrator_name = lower(integrator["type"].is_null() ? std::string("path") : (std::string)integrator["type"]);
	if (!integrator["path_length"].is_null()) config.path_length = integrator["path_length"];
	if (!integrator["sky_col"].is_null()) config.sky_col = jv3(integrator["sky_col"]);

	const std::string mesh_file = root + std::string(j["mesh_file"]);
	tinyobj::ObjReaderConfig reader_config;
	tinyobj::ObjReader reader;
	if (!reader.ParseFromFile(mesh_file, reader_config)) throw std::runtime_error("TinyObjReader: " + reader.Error() + " (" + mesh_file + ")");
	const auto& attrib = reader.GetAttrib();
	const auto& shapes = reader.GetShapes();

	prim_meshes.resize(shapes.size());
	for (uint32_t s = 0; s < shapes.size(); s++) {
		append_obj_shape(&attrib, &shapes[s], prim_meshes[s]);
		prim_meshes[s].prim_idx = s;
		prim_meshes[s].world_matrix = glm::mat4(1);  // LumenScene.cpp:328 ("TODO: Implement world transforms")
	}

	auto& bsdfs_arr = j["bsdfs"];
	auto& lights_arr = j["lights"];
	materials.assign(bsdfs_arr.size(), lmb_material{});  // value-initialised: unspecified fields are 0
	lights.assign(lights_arr.size(), AnalyticLight{});
	int bsdf_idx = 0;
	for (auto& bsdf : bsdfs_arr) {
		lmb_material& mat = materials[bsdf_idx];
		mat.texture_id = -1;
		if (!bsdf["texture"].is_null()) {
			texture_paths.push_back(root + (std::string)bsdf["texture"]);
			mat.texture_id = (int)texture_paths.size() - 1;
		}
		put3(mat.albedo, bsdf["albedo"].is_null() ? glm::vec3(1) : jv3(bsdf["albedo"]));
		if (!bsdf["emissive_factor"].is_null()) put3(mat.emissive_factor, jv3(bsdf["emissive_factor"]));

		const std::string type = bsdf["type"].is_null() ? std::string() : (std::string)bsdf["type"];
		if (type == "diffuse") {
			bsdf_types |= LMB_BSDF_DIFFUSE;
			mat.bsdf_type = LMB_BSDF_DIFFUSE;
			mat.bsdf_props = LMB_FLAG_DIFFUSE | LMB_FLAG_REFLECTION;
		} else if (type == "mirror") {
			bsdf_types |= LMB_BSDF_MIRROR;
			mat.bsdf_type = LMB_BSDF_MIRROR;
			mat.bsdf_props = LMB_FLAG_SPECULAR | LMB_FLAG_REFLECTION;
		} else if (type == "glass") {
			bsdf_types |= LMB_BSDF_GLASS;
			mat.bsdf_type = LMB_BSDF_GLASS;
			mat.bsdf_props = LMB_FLAG_SPECULAR | LMB_FLAG_TRANSMISSION;
			mat.ior = bsdf["ior"];
		} else if (type == "dielectric") {
			bsdf_types |= LMB_BSDF_DIELECTRIC;
			mat.bsdf_type = LMB_BSDF_DIELECTRIC;
			mat.ior = jf(bsdf, "ior", 1.0f);
			mat.roughness = jf(bsdf, "roughness", 0.0f);
			const auto transmission = bsdf["transmission"], reflection = bsdf["reflection"];
			if (transmission.is_null() || bool(transmission)) mat.bsdf_props |= LMB_FLAG_TRANSMISSION;
			if (reflection.is_null() || bool(reflection)) mat.bsdf_props |= LMB_FLAG_REFLECTION;
			if (mat.ior != 1.0 && mat.roughness > 0.08)
				mat.bsdf_props |= LMB_FLAG_GLOSSY;
			else
				mat.bsdf_props |= LMB_FLAG_SPECULAR;
			mat.thin = ju(bsdf, "thin", 0);
		} else if (type == "conductor") {
			bsdf_types |= LMB_BSDF_CONDUCTOR;
			mat.bsdf_type = LMB_BSDF_CONDUCTOR;
			mat.roughness = jf(bsdf, "roughness", 0.0f);
			// albedo carries eta, k the absorption coefficient (LumenScene.cpp:413-422)
			if (!bsdf["reflectance"].is_null()) {
				const glm::vec3 r = glm::clamp(jv3(bsdf["reflectance"]), 0.0f, 0.9999f);
				glm::vec3 eta, k;
				reflectance_to_conductor_eta_k(r, eta, k);
				put3(mat.albedo, eta);
				put3(mat.k, k);
			}
			if (!bsdf["edge_tint"].is_null()) {  // Gulbrandsen 2014 mapping, LumenScene.cpp:424-435
				const glm::vec3 g = jv3(bsdf["edge_tint"]);
				const glm::vec3 r = jv3(bsdf["reflectivity"]);
				const glm::vec3 n = g * (1.0f - r) / (1.0f + r) + (1.0f - g) * (1.0f + glm::sqrt(r)) / (1.0f - glm::sqrt(r));
				const glm::vec3 t1 = r * (n + 1.0f);
				const glm::vec3 t2 = n - 1.0f;
				put3(mat.albedo, n);
				put3(mat.k, glm::sqrt(1.0f / (1.0f - r) * (t1 * t1 - t2 * t2)));
			}
			mat.bsdf_props = LMB_FLAG_REFLECTION;
			mat.bsdf_props |= (mat.roughness > 0.08) ? LMB_FLAG_GLOSSY : LMB_FLAG_SPECULAR;
		} else if (type == "principled") {
			bsdf_types |= LMB_BSDF_PRINCIPLED;
			mat.bsdf_type = LMB_BSDF_PRINCIPLED;
			put3(mat.albedo, bsdf["albedo"].is_null() ? glm::vec3(1) : jv3(bsdf["albedo"]));
			mat.ior = jf(bsdf, "ior", 1.0f);
			mat.roughness = jf(bsdf, "roughness", 0.5f);
			mat.diffuse_trans = jf(bsdf, "diffuse_transmission", 0.0f);
			mat.spec_trans = jf(bsdf, "specular_transmission", 0.0f);
			mat.metallic = jf(bsdf, "metallic", 0.0f);
			mat.specular_tint = jf(bsdf, "specular_tint", 0.0f);
			mat.sheen_tint = jf(bsdf, "sheen_tint", 0.5f);
			mat.clearcoat = jf(bsdf, "clearcoat", 0.0f);
			mat.clearcoat_gloss = jf(bsdf, "clearcoat_gloss", 1.0f);
			mat.subsurface = jf(bsdf, "subsurface", 0.0f);
			mat.flatness = jf(bsdf, "flatness", 0.0f);
			mat.sheen = jf(bsdf, "sheen", 0.0f);
			mat.anisotropy = jf(bsdf, "anisotropy", 0.0f);
			mat.thin = ju(bsdf, "thin", 0);
			if (mat.roughness < 1.0f) mat.bsdf_props |= LMB_FLAG_REFLECTION;
			if (mat.spec_trans > 0.0f) mat.bsdf_props |= LMB_FLAG_TRANSMISSION;
			mat.bsdf_props |= (mat.roughness > 0.08) ? LMB_FLAG_GLOSSY : LMB_FLAG_SPECULAR;
		}
		// any other type string (e.g. "disney" in cornell_box_disney.json) leaves bsdf_type 0: paths die there (Q8)

		for (auto& ref : bsdf["refs"])
			for (size_t s = 0; s < shapes.size(); s++)
				if (ref == shapes[s].name) prim_meshes[s].material_idx = bsdf_idx;
		bsdf_idx++;
	}

	config.cam.fov = j["camera"]["fov"];
	config.cam.pos = jv3(j["camera"]["position"]);
	config.cam.dir = jv3(j["camera"]["dir"]);
	compute_scene_dimensions();
	int light_idx = 0;
	for (auto& light : lights_arr) {
		AnalyticLight& l = lights[light_idx++];
		l.pos = jv3(light["pos"]);
		l.to = jv3(light["dir"]);  // a point, not a direction: light_dir = normalize(to - pos)
		l.L = jv3(light["L"]);
		if (light["type"] == "spot")
			l.light_flags |= LMB_LIGHT_SPOT | LMB_LIGHT_FINITE_BIT | LMB_LIGHT_DELTA_BIT;
		else if (light["type"] == "directional")
			l.light_flags |= LMB_LIGHT_DIRECTIONAL | LMB_LIGHT_DELTA_BIT;
	}
}

// ------------------------------------------------------------------------------------------------ Mitsuba XML
namespace {
struct MBsdf {
	std::string name, type, texture;
	glm::vec3 albedo{1};
	float roughness = 0, ior = 1.0f;
};
struct MMesh {
	std::string file, shape_type;
	int bsdf_idx = -1;
	int inline_bsdf = -1;  // index into the loader's list of id-less <bsdf> children (opt-in emitters only)
	glm::mat4 transform{1};
	bool has_emitter = false;  // nested <emitter type="area">
	glm::vec3 radiance{0};
};
struct MLight {
	std::string type;
	glm::vec3 from{0}, to{0}, L{0};
};
}  // namespace

void Scene::load_mitsuba_scene(const std::string& path) {
	using namespace TPM_NAMESPACE;
	const std::string root = path.substr(0, path.find_last_of("/\\") + 1);
	SceneLoader loader;
	auto scene = loader.loadFromFile(path);

	// MitsubaParser::parse, MitsubaParser.cpp:5-147
	std::vector<MBsdf> bsdfs;
	std::vector<MMesh> meshes;
	std::vector<MLight> mlights;
	std::string integrator_type = "path";
	int depth = config.path_length;
	glm::vec3 sky(0.0f);
	float cam_fov = 45.0f;
	glm::mat4 cam_matrix(0.0f);
	// MitsubaParser.cpp:40-73
	auto parse_bsdf = [](Object* obj) {
		MBsdf b;
		b.name = obj->id();
		while ((obj->pluginType() == "twosided" || obj->pluginType() == "mask") && obj->anonymousChildren().size())
			obj = obj->anonymousChildren()[0].get();
		b.type = obj->pluginType();
		for (const auto& prop : obj->properties()) {
			if (prop.second.type() == PT_COLOR) {
				if (prop.first.find("reflectance") == std::string::npos && prop.first.find("specularReflectance") == std::string::npos)
					continue;
				b.albedo = glm::vec3((float)prop.second.getColor().r, (float)prop.second.getColor().g, (float)prop.second.getColor().b);
			}
			if (prop.first == "alpha") b.roughness = std::sqrt((float)prop.second.getNumber());
			if (prop.first == "int_ior") b.ior = (float)prop.second.getNumber();
		}
		for (const auto& nc : obj->namedChildren())
			if (nc.second->type() == OT_TEXTURE)
				for (const auto& tp : nc.second->properties())
					if (tp.first == "filename") b.texture = tp.second.getString();
		return b;
	};
	std::vector<MBsdf> inline_bsdfs;  // <bsdf> nested in a shape without an id (bedroom's lamp rectangles); only the opt-in emitters use them
	for (const auto& child : scene.anonymousChildren()) {
		Object* obj = child.get();
		switch (obj->type()) {
			case OT_INTEGRATOR: {
				integrator_type = obj->pluginType();
				for (const auto& prop : obj->properties())
					if (prop.first == "max_depth") depth = (int)prop.second.getInteger();
			} break;
			case OT_SENSOR: {
				for (const auto& prop : obj->properties()) {
					if (prop.first == "fov") {
						cam_fov = (float)prop.second.getNumber();
					} else if (prop.first == "to_world") {
						// straight copy of the row-major source into glm's column-major storage (MitsubaParser.cpp:30-37)
						float* dst = glm::value_ptr(cam_matrix);
						const auto& src = prop.second.getTransform();
						for (int i = 0; i < 16; i++) dst[i] = (float)src.matrix[i];
					}
				}
			} break;
			case OT_BSDF: {
				bsdfs.push_back(parse_bsdf(obj));
			} break;
			case OT_SHAPE: {
				MMesh m;
				m.shape_type = obj->pluginType();
				for (const auto& mc : obj->anonymousChildren()) {
					if (mc->type() != OT_EMITTER || mc->pluginType() != "area") continue;
					for (const auto& ep : mc->properties()) {
						if (ep.first != "radiance") continue;
						m.has_emitter = true;
						if (ep.second.type() == PT_COLOR)
							m.radiance = glm::vec3((float)ep.second.getColor().r, (float)ep.second.getColor().g, (float)ep.second.getColor().b);
						else if (ep.second.type() == PT_NUMBER || ep.second.type() == PT_INTEGER)
							m.radiance = glm::vec3((float)ep.second.getNumber());
					}
				}
				for (const auto& prop : obj->properties()) {
					if (prop.first == "filename") {
						m.file = prop.second.getString();
					} else if (prop.first == "to_world") {
						float* dst = glm::value_ptr(m.transform);  // transposed copy (MitsubaParser.cpp:98-106)
						const auto& src = prop.second.getTransform();
						for (int i = 0; i < 4; i++)
							for (int k = 0; k < 4; k++) dst[4 * i + k] = (float)src.matrix[4 * k + i];
					}
				}
				for (const auto& mc : obj->anonymousChildren()) {
					const auto ref = mc->id();
					for (size_t i = 0; i < bsdfs.size(); i++)
						if (bsdfs[i].name == ref) m.bsdf_idx = (int)i;
					if (mc->type() == OT_BSDF && ref.empty()) {
						inline_bsdfs.push_back(parse_bsdf(mc.get()));
						m.inline_bsdf = (int)inline_bsdfs.size() - 1;
					}
				}
				meshes.push_back(m);
			} break;
			case OT_EMITTER: {
				// MitsubaParser.cpp:121-142. The reference folds sun_scale into L while iterating an unordered_map, so its
				// result depends on hash order; this build fixes the intended value L = sun_color * sun_scale.
				MLight l;
				if (obj->pluginType() == "sunsky") l.type = "directional";
				glm::vec3 color(0.0f);
				float scale = 1.0f;
				for (const auto& prop : obj->properties()) {
					if (prop.first == "sun_direction") {
						const auto d = prop.second.getVector();
						l.from = glm::vec3((float)d.x, (float)d.y, (float)d.z);
					} else if (prop.first == "sun_color") {
						const auto c = prop.second.getVector();
						color = glm::vec3((float)c.x, (float)c.y, (float)c.z);
					} else if (prop.first == "sun_scale") {
						scale = (float)prop.second.getNumber();
					} else if (prop.first == "sky_color") {
						const auto c = prop.second.getVector();
						sky = glm::vec3((float)c.x, (float)c.y, (float)c.z);
					}
				}
				l.L = color * scale;
				mlights.push_back(l);
			} break;
			default:
				break;
		}
	}

	// LumenScene::load_mitsuba_scene, LumenScene.cpp:514-690
	config.integrator_name = lower(integrator_type);
	config.path_length = depth;
	config.sky_col = sky;
	config.cam.fov = cam_fov / 2;  // LumenScene.cpp:532
	config.cam.cam_matrix = cam_matrix;
	config.cam.pos = glm::vec3(0);
	// Q11: shapes without a file (rectangle emitters etc.) are skipped. The reference keeps zero-sized trailing
	// prim-mesh slots for them; they hold no triangles and are dropped here.
	// SURVEY.md 8f-2, OPT-IN (LUMEN_B200_MITSUBA_AREA_EMITTERS=1; off by default because the reference drops them, LumenScene.cpp:
	// 538-540, MitsubaParser.cpp:121-142): `rectangle` shapes become two triangles (Mitsuba's [-1,1]^2 in the XY plane, normal +Z)
	// and a nested <emitter type="area"> makes the shape's material emissive, i.e. an area light of Lumen's own kind
	// (LumenScene.cpp:83-95). Emitter geometry is baked to world space with an identity world matrix: sample_triangle puts
	// w = 1 on edge vectors (quirk Q7), which is only right without a translation.
	const char* opt = std::getenv("LUMEN_B200_MITSUBA_AREA_EMITTERS");
	const bool area_emitters = opt && *opt && std::strcmp(opt, "0") != 0;
	std::vector<std::pair<size_t, glm::vec3>> emissive;  // (material index, radiance)
	auto bake_to_world = [&](PrimMesh& pm) {
		const glm::mat3 nm = glm::transpose(glm::inverse(glm::mat3(pm.world_matrix)));
		glm::vec3 mn(FLT_MAX), mx(-FLT_MAX);
		for (uint32_t v = pm.vtx_offset; v < pm.vtx_offset + pm.idx_count; v++) {
			positions[v] = glm::vec3(pm.world_matrix * glm::vec4(positions[v], 1.0f));
			normals[v] = glm::normalize(nm * normals[v]);
			mn = glm::min(mn, positions[v]), mx = glm::max(mx, positions[v]);
		}
		pm.min_pos = mn, pm.max_pos = mx;
		pm.world_matrix = glm::mat4(1.0f);
	};
	auto emitter_material = [&](const MMesh& mesh) -> uint32_t {
		MBsdf b = mesh.bsdf_idx >= 0 ? bsdfs[(size_t)mesh.bsdf_idx] : (mesh.inline_bsdf >= 0 ? inline_bsdfs[(size_t)mesh.inline_bsdf] : MBsdf{});
		if (mesh.bsdf_idx < 0 && mesh.inline_bsdf < 0) b.type = "diffuse", b.albedo = glm::vec3(0.5f);  // Mitsuba's default bsdf
		b.name += "#emitter" + std::to_string(emissive.size());
		bsdfs.push_back(b);
		emissive.emplace_back(bsdfs.size() - 1, mesh.radiance);
		return (uint32_t)(bsdfs.size() - 1);
	};
	for (const auto& mesh : meshes) {
		if (mesh.file.empty()) {
			if (!area_emitters || mesh.shape_type != "rectangle") continue;
			PrimMesh pm;
			pm.name = "rectangle";
			pm.first_idx = (uint32_t)indices.size();
			pm.vtx_offset = (uint32_t)positions.size();
			pm.idx_count = 6, pm.vtx_count = 2;
			static const float q[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
			static const int tri[6] = {0, 1, 2, 0, 2, 3};
			for (uint32_t k = 0; k < 6; k++) {
				indices.push_back(k);
				positions.emplace_back(q[tri[k]][0], q[tri[k]][1], 0.0f);
				normals.emplace_back(0.0f, 0.0f, 1.0f);
				texcoords0.emplace_back(0.5f * (q[tri[k]][0] + 1.0f), 0.5f * (q[tri[k]][1] + 1.0f));
			}
			pm.prim_idx = (uint32_t)prim_meshes.size();
			pm.world_matrix = mesh.transform;
			bake_to_world(pm);
			if (mesh.has_emitter)
				pm.material_idx = emitter_material(mesh);
			else if (mesh.bsdf_idx >= 0)
				pm.material_idx = (uint32_t)mesh.bsdf_idx;
			else
				throw std::runtime_error("rectangle shape without bsdf ref or area emitter");
			prim_meshes.push_back(pm);
			continue;
		}
		const std::string mesh_file = root + mesh.file;
		tinyobj::ObjReaderConfig reader_config;
		tinyobj::ObjReader reader;
		if (!reader.ParseFromFile(mesh_file, reader_config)) throw std::runtime_error("TinyObjReader: " + reader.Error() + " (" + mesh_file + ")");
		const auto& shapes = reader.GetShapes();
		if (shapes.size() != 1) throw std::runtime_error("Mitsuba OBJ must hold exactly one shape: " + mesh_file);
		PrimMesh pm;
		append_obj_shape(&reader.GetAttrib(), &shapes[0], pm);
		pm.prim_idx = (uint32_t)prim_meshes.size();
		pm.world_matrix = mesh.transform;
		if (area_emitters && mesh.has_emitter) {
			bake_to_world(pm);
			pm.material_idx = emitter_material(mesh);
		} else {
			if (mesh.bsdf_idx < 0) throw std::runtime_error("shape without a resolvable bsdf ref: " + mesh_file);
			pm.material_idx = (uint32_t)mesh.bsdf_idx;
		}
		prim_meshes.push_back(pm);
	}

	materials.assign(bsdfs.size(), lmb_material{});
	for (size_t i = 0; i < bsdfs.size(); i++) {
		const MBsdf& b = bsdfs[i];
		lmb_material& mat = materials[i];
		if (!b.texture.empty()) {
			texture_paths.push_back(root + b.texture);
			mat.texture_id = (int)texture_paths.size() - 1;
		} else {
			mat.texture_id = -1;
		}
		put3(mat.albedo, b.albedo);
		mat.roughness = b.roughness;
		if (b.type == "diffuse") {
			bsdf_types |= LMB_BSDF_DIFFUSE;
			mat.bsdf_type = LMB_BSDF_DIFFUSE;
			mat.bsdf_props = LMB_FLAG_DIFFUSE | LMB_FLAG_REFLECTION;
		} else if (b.type == "roughplastic" || b.type == "roughdielectric" || b.type == "dielectric" || b.type == "plastic") {
			bsdf_types |= LMB_BSDF_PRINCIPLED;
			mat.bsdf_type = LMB_BSDF_PRINCIPLED;
			mat.ior = b.ior;
			if (mat.roughness < 1.0) mat.bsdf_props |= LMB_FLAG_DIFFUSE | LMB_FLAG_REFLECTION;
			if (mat.ior != 1.0) mat.bsdf_props |= LMB_FLAG_TRANSMISSION;
			mat.bsdf_props |= (mat.roughness > 0.08) ? LMB_FLAG_GLOSSY : LMB_FLAG_SPECULAR;
			if (b.type == "roughdielectric" || b.type == "dielectric") {
				mat.spec_trans = 1.0;
				mat.metallic = 0.0;
			}
			if (b.type == "roughplastic" || b.type == "plastic") {
				mat.metallic = 1.0;
				mat.subsurface = 0.1f;
				mat.spec_trans = 0.5;
				mat.thin = 1;
			}
		} else if (b.type == "conductor" || b.type == "roughconductor") {
			bsdf_types |= LMB_BSDF_CONDUCTOR;
			mat.bsdf_type = LMB_BSDF_CONDUCTOR;
			glm::vec3 eta, k;
			reflectance_to_conductor_eta_k(b.albedo, eta, k);
			put3(mat.albedo, eta);
			put3(mat.k, k);
			mat.bsdf_props = LMB_FLAG_REFLECTION;
			mat.bsdf_props |= (mat.roughness > 0.08) ? LMB_FLAG_GLOSSY : LMB_FLAG_SPECULAR;
		} else if (b.type == "glass") {
			bsdf_types |= LMB_BSDF_GLASS;
			mat.bsdf_type = LMB_BSDF_GLASS;
			mat.bsdf_props = LMB_FLAG_SPECULAR | LMB_FLAG_TRANSMISSION;
			mat.ior = b.ior;
		}
	}
	for (const auto& e : emissive) put3(materials[e.first].emissive_factor, e.second);
	compute_scene_dimensions();
	lights.assign(mlights.size(), AnalyticLight{});
	for (size_t i = 0; i < mlights.size(); i++) {
		lights[i].L = 100.0f * mlights[i].L;
		if (mlights[i].type == "directional") {
			lights[i].pos = mlights[i].from;
			lights[i].to = mlights[i].to;
			lights[i].light_flags = LMB_LIGHT_DIRECTIONAL | LMB_LIGHT_DELTA_BIT;
		}
	}
}

// ------------------------------------------------------------------------------------------------ common
void Scene::compute_scene_dimensions() {
	Bbox scene_bbox;
	for (const auto& pm : prim_meshes) {
		// LumenScene.cpp:735-737: `bbox.transform(pm.world_matrix)` returns a new box that is discarded, so the
		// OBJECT-space bounds are what gets inserted. Kept as is (identical for identity transforms).
		scene_bbox.insert(pm.min_pos);
		scene_bbox.insert(pm.max_pos);
	}
	if (scene_bbox.is_empty() || !scene_bbox.is_volume()) {
		scene_bbox.insert({-1.0f, -1.0f, -1.0f});
		scene_bbox.insert({1.0f, 1.0f, 1.0f});
	}
	dim_min = scene_bbox.mn;
	dim_max = scene_bbox.mx;
	dim_radius = glm::length(scene_bbox.mx - scene_bbox.mn) * 0.5f;
}

void Scene::load(const std::string& path, uint32_t w, uint32_t h) {
	if (ends_with(path, ".json"))
		load_lumen_scene(path);
	else if (ends_with(path, ".xml"))
		load_mitsuba_scene(path);
	else
		throw std::runtime_error("unknown scene format: " + path);
	finalize(w, h);
}

void Scene::finalize(uint32_t w, uint32_t h) {
	width = w, height = h;
	const float aspect_ratio = (float)w / (float)h;
	if (config.cam.pos != glm::vec3(0))
		camera.init_lookat(config.cam.fov, aspect_ratio, config.cam.dir, config.cam.pos);
	else
		camera.init_matrix(config.cam.fov, config.cam.cam_matrix, aspect_ratio);

	// LumenScene.cpp:70-97
	total_light_triangle_cnt = 0;
	total_light_area = 0;
	prim_lookup.clear(), gpu_lights.clear(), prim_idx_counts.clear(), world_matrices.clear(), inv_world_matrices.clear();
	uint32_t idx = 0;
	for (auto& pm : prim_meshes) {
		if (pm.material_idx >= materials.size()) throw std::runtime_error("prim mesh '" + pm.name + "' has no material");
		lmb_prim_mesh_info info{};
		info.index_offset = pm.first_idx;
		info.vertex_offset = pm.vtx_offset;
		info.material_index = pm.material_idx;
		put3(info.min_pos, pm.min_pos);
		put3(info.max_pos, pm.max_pos);
		prim_lookup.push_back(info);
		prim_idx_counts.push_back(pm.idx_count);
		const glm::mat4 inv = glm::inverse(pm.world_matrix);
		world_matrices.insert(world_matrices.end(), glm::value_ptr(pm.world_matrix), glm::value_ptr(pm.world_matrix) + 16);
		inv_world_matrices.insert(inv_world_matrices.end(), glm::value_ptr(inv), glm::value_ptr(inv) + 16);
		const glm::vec3 mef = get3(materials[pm.material_idx].emissive_factor);
		if (mef.x > 0 || mef.y > 0 || mef.z > 0) {
			lmb_light light{};
			std::memcpy(light.world_matrix, glm::value_ptr(pm.world_matrix), 64);
			light.num_triangles = pm.idx_count / 3;
			light.prim_mesh_idx = idx;
			light.light_flags = LMB_LIGHT_AREA | LMB_LIGHT_FINITE_BIT;
			put3(light.L, mef);
			gpu_lights.push_back(light);
			total_light_triangle_cnt += light.num_triangles;
		}
		idx++;
	}
	// LumenScene.cpp:99-113
	for (size_t i = 0; i < lights.size(); i++) {
		const AnalyticLight& l = lights[i];
		lmb_light light{};
		put3(light.L, l.L);
		light.light_flags = l.light_flags;
		put3(light.pos, l.pos);
		put3(light.to, l.to);
		total_light_triangle_cnt++;
		light.world_radius = dim_radius;
		put3(light.world_center, 0.5f * (dim_max + dim_min));
		if ((l.light_flags & LMB_LIGHT_DIRECTIONAL) == LMB_LIGHT_DIRECTIONAL) dir_light_idx = (uint32_t)i;  // Q6: index into `lights`
		gpu_lights.push_back(light);
	}
	// LumenScene.cpp:115-133
	float area_sum = 0.0f;
	for (auto& l : gpu_lights) {
		if ((l.light_flags & 0x7) != LMB_LIGHT_AREA) continue;
		const PrimMesh& pm = prim_meshes[l.prim_mesh_idx];
		for (uint32_t i = 0; i < l.num_triangles; i++) {
			const uint32_t o = pm.first_idx + 3 * i;
			const glm::vec3 v0 = pm.world_matrix * glm::vec4(positions[indices[o] + pm.vtx_offset], 1.0);
			const glm::vec3 v1 = pm.world_matrix * glm::vec4(positions[indices[o + 1] + pm.vtx_offset], 1.0);
			const glm::vec3 v2 = pm.world_matrix * glm::vec4(positions[indices[o + 2] + pm.vtx_offset], 1.0);
			area_sum += 0.5f * glm::length(glm::cross(v1 - v0, v2 - v0));
		}
	}
	total_light_area += area_sum;
	// LumenScene.cpp:174-184
	vertices.resize(positions.size());
	for (size_t i = 0; i < positions.size(); i++) {
		put3(vertices[i].pos, positions[i]);
		put3(vertices[i].normal, normals[i]);
		vertices[i].uv0[0] = texcoords0[i].x, vertices[i].uv0[1] = texcoords0[i].y;
	}
	// LumenScene.cpp:200-216: textures decoded to 4 x 8 bit
	texture_data.clear();
	texture_views.clear();
	for (const auto& tp : texture_paths) {
		int x, y, n;
		unsigned char* data = stbi_load(tp.c_str(), &x, &y, &n, 4);
		if (!data) throw std::runtime_error("cannot load texture " + tp);
		Texture t;
		t.w = (uint32_t)x, t.h = (uint32_t)y;
		t.rgba8.assign(data, data + (size_t)x * y * 4);
		stbi_image_free(data);
		texture_data.push_back(std::move(t));
	}
	for (const auto& t : texture_data) texture_views.push_back(lmb_texture{t.rgba8.data(), t.w, t.h});
}

lmb_scene_desc Scene::desc() const {
	lmb_scene_desc d{};
	d.vertices = vertices.data(), d.n_vertices = (uint32_t)vertices.size();
	d.indices = indices.data(), d.n_indices = (uint32_t)indices.size();
	d.materials = materials.data(), d.n_materials = (uint32_t)materials.size();
	d.prim_infos = prim_lookup.data(), d.n_prim_meshes = (uint32_t)prim_lookup.size();
	d.prim_idx_counts = prim_idx_counts.data();
	d.world_matrices = world_matrices.data();
	d.inv_world_matrices = inv_world_matrices.data();
	d.lights = gpu_lights.data(), d.n_lights = (uint32_t)gpu_lights.size();
	d.textures = texture_views.data(), d.n_textures = (uint32_t)texture_views.size();
	return d;
}

lmb_pc_path Scene::make_pc(int max_depth, bool direct_lighting) const {
	lmb_pc_path pc{};
	pc.size_x = width, pc.size_y = height;
	pc.num_lights = (int)gpu_lights.size();
	pc.time = 0;
	pc.max_depth = max_depth > 0 ? max_depth : config.path_length;
	put3(pc.sky_col, config.sky_col);
	pc.total_light_area = total_light_area;
	pc.light_triangle_count = (int)total_light_triangle_cnt;
	pc.dir_light_idx = dir_light_idx;
	pc.frame_num = 0;
	pc.direct_lighting = direct_lighting ? 1u : 0u;
	return pc;
}

lmb_scene_ubo Scene::make_ubo() {
	camera.update_view_matrix();
	lmb_scene_ubo u{};
	auto put = [](float* dst, const glm::mat4& m) { std::memcpy(dst, glm::value_ptr(m), 64); };
	put(u.prev_view, camera.view);
	put(u.view, camera.view);
	put(u.prev_projection, camera.projection);
	put(u.projection, camera.projection);
	u.view_pos[0] = camera.position.x, u.view_pos[1] = camera.position.y, u.view_pos[2] = camera.position.z, u.view_pos[3] = 1;
	put(u.inv_view, glm::inverse(camera.view));
	put(u.inv_projection, glm::inverse(camera.projection));
	put(u.model, glm::mat4(1.0f));
	u.light_pos[0] = 3.0f, u.light_pos[1] = 2.5f, u.light_pos[2] = 1.0f, u.light_pos[3] = 1.0f;
	return u;
}

}  // namespace lmh
