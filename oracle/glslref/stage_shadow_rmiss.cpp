// ORACLE -- TEST INFRASTRUCTURE ONLY. src/shaders/ray_shadow.rmiss (miss index 1), translated by glsl2cpp.py.
#include "stage_common.h"
namespace glslref {
struct ShadowRmissStage : Stage {
	using Stage::Stage;
#include "gen/ray_shadow.rmiss.inc"
};
void run_shadow_rmiss(const Stage::Inputs& in) {
	ShadowRmissStage st(in);
	st.main();
}
}  // namespace glslref
