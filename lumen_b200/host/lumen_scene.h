// Host-side scene ingest for the B200 path integrator: the un-Vulkan'd equivalent of Lumen's LumenScene
// (reference: src/RayTracer/LumenScene.{h,cpp}, src/RayTracer/SceneConfig.h, src/Framework/MitsubaParser.{h,cpp},
// src/Framework/Camera.h, src/Framework/BBox.h). It turns a Lumen JSON scene (+OBJ) or a Mitsuba XML scene (+OBJs)
// into exactly the flat arrays Lumen uploads for `Path` (LumenScene.cpp:134-216) and exposes them as an
// lmb_scene_desc for lmb_upload_scene().
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include <glm/glm.hpp>

#include "lmb_types.h"

namespace lmh {

struct PrimMesh {  // LumenScene.h:27-38 (LumenPrimMesh)
	std::string name;
	uint32_t material_idx = 0;
	uint32_t vtx_offset = 0;
	uint32_t first_idx = 0;
	uint32_t idx_count = 0;
	uint32_t vtx_count = 0;
	uint32_t prim_idx = 0;
	glm::mat4 world_matrix{1.0f};
	glm::vec3 min_pos{0.0f};
	glm::vec3 max_pos{0.0f};
};

struct AnalyticLight {  // LumenScene.h:40-47 (LumenLight)
	glm::vec3 pos{0.0f};
	glm::vec3 to{0.0f};
	glm::vec3 L{0.0f};
	uint32_t light_flags = 0;
};

struct CameraSettings {  // SceneConfig.h:6-11
	float fov = 45.0f;
	glm::vec3 pos{0.0f};
	glm::vec3 dir{0.0f};
	glm::mat4 cam_matrix{0.0f};
};

struct SceneConfig {  // SceneConfig.h:15-27
	int path_length = 6;
	glm::vec3 sky_col{0.0f};
	std::string integrator_name = "path";  // lower-cased `integrator.type`; unknown names fall back to path
	CameraSettings cam;
};

// Perspective camera, Camera.h:49-126. The view matrix Lumen renders with is rebuilt every frame from `position` and
// the Euler `rotation` (degrees) as inverse(T * Ry * Rx * Rz) (Camera.h:28-39); the constructors only seed `rotation`.
struct PerspectiveCamera {
	glm::mat4 projection{1.0f};
	glm::mat4 view{1.0f};
	glm::vec3 position{0.0f};
	glm::vec3 rotation{0.0f};
	float fov = 0, aspect_ratio = 1, cam_near = 0.01f, cam_far = 1000.0f;
	void init_lookat(float fov_deg, float aspect, const glm::vec3& dir, const glm::vec3& pos);  // Camera.h:67-85
	void init_matrix(float fov_deg, const glm::mat4& cam_matrix, float aspect);                   // Camera.h:87-105
	void update_view_matrix();                                                                  // Camera.h:28-39
  private:
	void make_projection();  // Camera.h:109-123 (use_fov == true)
};

class Scene {
  public:
	// Image size replaces Lumen's Window::width()/height().
	void load(const std::string& path, uint32_t width, uint32_t height);
	// Post-processing shared by every ingest route (LumenScene.cpp:59-190): camera, PrimMeshInfo table, area-light
	// discovery, analytic lights, light totals, interleaved Vertex array.
	void finalize(uint32_t width, uint32_t height);

	// raw geometry (de-indexed: 3 vertices per triangle, LumenScene.cpp:302-325)
	std::vector<glm::vec3> positions, normals;
	std::vector<glm::vec2> texcoords0;
	std::vector<uint32_t> indices;
	std::vector<PrimMesh> prim_meshes;
	std::vector<lmb_material> materials;
	std::vector<std::string> texture_paths;
	std::vector<AnalyticLight> lights;
	SceneConfig config;

	// derived
	std::vector<lmb_vertex> vertices;
	std::vector<lmb_prim_mesh_info> prim_lookup;
	std::vector<uint32_t> prim_idx_counts;
	std::vector<float> world_matrices, inv_world_matrices;
	std::vector<lmb_light> gpu_lights;
	struct Texture {
		std::vector<uint8_t> rgba8;
		uint32_t w = 0, h = 0;
	};
	std::vector<Texture> texture_data;
	std::vector<lmb_texture> texture_views;
	PerspectiveCamera camera;
	uint32_t total_light_triangle_cnt = 0;
	float total_light_area = 0;
	uint32_t dir_light_idx = 0xFFFFFFFFu;
	uint32_t bsdf_types = 0;
	glm::vec3 dim_min{0}, dim_max{0};
	float dim_radius = 0;
	uint32_t width = 0, height = 0;

	lmb_scene_desc desc() const;
	// Path::render's push-constant fill (Path.cpp:27-38). frame_num is left 0; `time` is 0 (unused by the shader, Q12).
	lmb_pc_path make_pc(int max_depth, bool direct_lighting) const;
	// Integrator::update_uniform_buffers (Integrator.cpp:60-72)
	lmb_scene_ubo make_ubo();

	void compute_scene_dimensions();  // LumenScene.cpp:732-751

  private:
	void load_lumen_scene(const std::string& path);    // LumenScene.cpp:231-513
	void load_mitsuba_scene(const std::string& path);  // LumenScene.cpp:514-690
	void append_obj_shape(const void* attrib, const void* shape, PrimMesh& pm);
};

// ImageUtils.cpp:22-89: RGBA fp32 -> 3-channel (B,G,R) HALF OpenEXR.
bool save_exr(const float* rgba, int width, int height, const char* path, std::string* err);
// ImageUtils.cpp:8-20
bool save_exr_half_bgr(const uint16_t* planes, int width, int height, const char* path, std::string* err);
// progressive-render checkpoint: header (magic, width, height, frames accumulated, path length) + RGBA32F film
bool save_checkpoint(const char* path, const float* rgba, uint32_t width, uint32_t height, uint32_t frames, uint32_t path_length, std::string* err);
bool load_checkpoint(const char* path, std::vector<float>& rgba, uint32_t& width, uint32_t& height, uint32_t& frames, uint32_t& path_length,
					 std::string* err);
bool load_exr(const char* path, std::vector<float>& rgba, int& width, int& height, std::string* err);

}  // namespace lmh
