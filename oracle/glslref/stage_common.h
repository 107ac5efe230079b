// ORACLE -- TEST INFRASTRUCTURE ONLY (see prelude.h).
// Shared by the stage translation units: the reference's host/device headers, included UNMODIFIED from the reference
// tree at namespace scope (they are C++ already: commons.h:38-57), and the layout cross-checks against include/lmb_types.h.
#pragma once
#include "prelude.h"
// reference headers, where they lie (-I$(REF)/src/shaders)
#include "commons.h"
#include "integrators/path/path_commons.h"
#include "lmb_types.h"

static_assert(sizeof(::Vertex) == sizeof(lmb_vertex), "Vertex: commons.h vs lmb_types.h");
static_assert(sizeof(::Light) == sizeof(lmb_light), "Light");
static_assert(sizeof(::Material) == sizeof(lmb_material), "Material");
static_assert(sizeof(::PrimMeshInfo) == sizeof(lmb_prim_mesh_info), "PrimMeshInfo");
static_assert(sizeof(::PCPath) == sizeof(lmb_pc_path), "PCPath");
static_assert(sizeof(::SceneUBO) == sizeof(lmb_scene_ubo), "SceneUBO");
static_assert(offsetof(::Material, texture_id) == offsetof(lmb_material, texture_id), "Material.texture_id");
static_assert(offsetof(::Material, thin) == offsetof(lmb_material, thin), "Material.thin");
static_assert(offsetof(::Light, light_flags) == offsetof(lmb_light, light_flags), "Light.light_flags");
static_assert(offsetof(::Light, world_radius) == offsetof(lmb_light, world_radius), "Light.world_radius");
static_assert(offsetof(::PCPath, dir_light_idx) == offsetof(lmb_pc_path, dir_light_idx), "PCPath.dir_light_idx");
static_assert(offsetof(::SceneUBO, inv_projection) == offsetof(lmb_scene_ubo, inv_projection), "SceneUBO.inv_projection");

// LumenScene.cpp:217-228 defines one ENABLE_* macro per BSDF type present in the scene; with all of them defined every
// `switch (mat.bsdf_type)` of bsdf_commons.glsl has all its cases, which is the same program for any scene.
#define ENABLE_DIFFUSE
#define ENABLE_MIRROR
#define ENABLE_GLASS
#define ENABLE_DIELECTRIC
#define ENABLE_CONDUCTOR
#define ENABLE_PRINCIPLED

namespace glslref {
// stage entry points used by the shader-binding-table emulation (harness.cpp)
void run_rchit(const Stage::Inputs& in);
void run_rmiss(const Stage::Inputs& in);
void run_shadow_rmiss(const Stage::Inputs& in);
}  // namespace glslref
