// C wrappers over lmh::Scene for ctypes and for the C++ headless renderer. See include/lumen_host.h.
#include "lumen_host.h"

#include <cstdlib>
#include <cstring>
#include <string>

#include <glm/gtc/type_ptr.hpp>

#include "lumen_scene.h"

struct lmh_scene {
	lmh::Scene scene;
};

static thread_local std::string g_err;

extern "C" {

const char* lmh_last_error(void) { return g_err.c_str(); }

int lmh_scene_load(const char* path, uint32_t width, uint32_t height, lmh_scene** out) {
	try {
		auto* s = new lmh_scene();
		try {
			s->scene.load(path, width, height);
		} catch (...) {
			delete s;
			throw;
		}
		*out = s;
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

int lmh_scene_from_arrays(const lmb_vertex* vertices, uint32_t n_vertices, const uint32_t* mesh_tri_counts, const uint32_t* mesh_materials,
						  const float* mesh_world, uint32_t n_meshes, const lmb_material* materials, uint32_t n_materials,
						  const lmb_light* analytic_lights, uint32_t n_analytic, float fov, const float* cam_pos, const float* cam_dir,
						  int32_t path_length, const float* sky_col, uint32_t width, uint32_t height, lmh_scene** out) {
	try {
		auto* h = new lmh_scene();
		lmh::Scene& sc = h->scene;
		sc.positions.resize(n_vertices), sc.normals.resize(n_vertices), sc.texcoords0.resize(n_vertices);
		for (uint32_t i = 0; i < n_vertices; i++) {
			sc.positions[i] = glm::make_vec3(vertices[i].pos);
			sc.normals[i] = glm::make_vec3(vertices[i].normal);
			sc.texcoords0[i] = glm::make_vec2(vertices[i].uv0);
		}
		uint32_t voff = 0;
		for (uint32_t m = 0; m < n_meshes; m++) {
			lmh::PrimMesh pm;
			pm.name = "mesh" + std::to_string(m);
			pm.material_idx = mesh_materials[m];
			pm.vtx_offset = voff;
			pm.first_idx = (uint32_t)sc.indices.size();
			pm.idx_count = 3 * mesh_tri_counts[m];
			pm.vtx_count = mesh_tri_counts[m];
			pm.prim_idx = m;
			pm.world_matrix = mesh_world ? glm::make_mat4(mesh_world + 16 * m) : glm::mat4(1.0f);
			glm::vec3 mn(3.402823466e+38f), mx(-3.402823466e+38f);
			for (uint32_t k = 0; k < pm.idx_count; k++) {
				sc.indices.push_back(k);
				mn = glm::min(mn, sc.positions[voff + k]);
				mx = glm::max(mx, sc.positions[voff + k]);
			}
			pm.min_pos = mn, pm.max_pos = mx;
			voff += pm.idx_count;
			sc.prim_meshes.push_back(pm);
		}
		if (voff != n_vertices) {
			delete h;
			g_err = "vertex count does not match 3 * sum(mesh_tri_counts)";
			return -1;
		}
		sc.materials.assign(materials, materials + n_materials);
		for (const auto& m : sc.materials) sc.bsdf_types |= m.bsdf_type;
		for (uint32_t i = 0; i < n_analytic; i++) {
			lmh::AnalyticLight l;
			l.pos = glm::make_vec3(analytic_lights[i].pos);
			l.to = glm::make_vec3(analytic_lights[i].to);
			l.L = glm::make_vec3(analytic_lights[i].L);
			l.light_flags = analytic_lights[i].light_flags;
			sc.lights.push_back(l);
		}
		sc.config.path_length = path_length;
		sc.config.sky_col = glm::make_vec3(sky_col);
		sc.config.cam.fov = fov;
		sc.config.cam.pos = glm::make_vec3(cam_pos);
		sc.config.cam.dir = glm::make_vec3(cam_dir);
		sc.compute_scene_dimensions();
		try {
			sc.finalize(width, height);
		} catch (...) {
			delete h;
			throw;
		}
		*out = h;
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

void lmh_scene_destroy(lmh_scene* s) { delete s; }
void lmh_scene_get_desc(const lmh_scene* s, lmb_scene_desc* out) { *out = s->scene.desc(); }
void lmh_scene_get_info(const lmh_scene* s, lmh_scene_info* out) {
	const lmh::Scene& sc = s->scene;
	std::memset(out, 0, sizeof(*out));
	out->path_length = sc.config.path_length;
	out->sky_col[0] = sc.config.sky_col.x, out->sky_col[1] = sc.config.sky_col.y, out->sky_col[2] = sc.config.sky_col.z;
	std::strncpy(out->integrator, sc.config.integrator_name.c_str(), sizeof(out->integrator) - 1);
	out->n_triangles = (uint32_t)(sc.indices.size() / 3);
	out->n_prim_meshes = (uint32_t)sc.prim_meshes.size();
	out->n_materials = (uint32_t)sc.materials.size();
	out->n_lights = (uint32_t)sc.gpu_lights.size();
	out->n_textures = (uint32_t)sc.texture_data.size();
	out->total_light_triangle_cnt = sc.total_light_triangle_cnt;
	out->total_light_area = sc.total_light_area;
	out->dir_light_idx = sc.dir_light_idx;
	out->bsdf_types = sc.bsdf_types;
	out->world_radius = sc.dim_radius;
}
void lmh_scene_make_pc(const lmh_scene* s, int32_t max_depth, int32_t direct_lighting, lmb_pc_path* out) {
	*out = s->scene.make_pc(max_depth, direct_lighting != 0);
}
void lmh_scene_make_ubo(lmh_scene* s, lmb_scene_ubo* out) { *out = s->scene.make_ubo(); }

int lmh_save_exr(const float* rgba, int32_t width, int32_t height, const char* path) {
	std::string err;
	if (!lmh::save_exr(rgba, width, height, path, &err)) {
		g_err = err;
		return -1;
	}
	return 0;
}
int lmh_save_exr_half_bgr(const uint16_t* planes, int32_t width, int32_t height, const char* path) {
	std::string err;
	if (!lmh::save_exr_half_bgr(planes, width, height, path, &err)) {
		g_err = err;
		return -1;
	}
	return 0;
}
int lmh_save_checkpoint(const char* path, const float* rgba, uint32_t width, uint32_t height, uint32_t frames, uint32_t path_length) {
	std::string err;
	if (!lmh::save_checkpoint(path, rgba, width, height, frames, path_length, &err)) {
		g_err = err;
		return -1;
	}
	return 0;
}
int lmh_load_checkpoint(const char* path, float** rgba_out, uint32_t* width, uint32_t* height, uint32_t* frames, uint32_t* path_length) {
	std::vector<float> px;
	std::string err;
	if (!lmh::load_checkpoint(path, px, *width, *height, *frames, *path_length, &err)) {
		g_err = err;
		return -1;
	}
	*rgba_out = (float*)std::malloc(px.size() * sizeof(float));
	std::memcpy(*rgba_out, px.data(), px.size() * sizeof(float));
	return 0;
}
int lmh_load_exr(const char* path, float** rgba_out, int32_t* width, int32_t* height) {
	std::vector<float> px;
	std::string err;
	int w = 0, h = 0;
	if (!lmh::load_exr(path, px, w, h, &err)) {
		g_err = err;
		return -1;
	}
	*rgba_out = (float*)std::malloc(px.size() * sizeof(float));
	std::memcpy(*rgba_out, px.data(), px.size() * sizeof(float));
	*width = w, *height = h;
	return 0;
}
void lmh_free(void* p) { std::free(p); }

}  // extern "C"
