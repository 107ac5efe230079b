// ORACLE -- TEST INFRASTRUCTURE ONLY. src/shaders/ray.rchit, translated by glsl2cpp.py: one invocation per closest hit.
#include "stage_common.h"
namespace glslref {
struct RchitStage : Stage {
	using Stage::Stage;
#include "gen/ray.rchit.inc"
};
void run_rchit(const Stage::Inputs& in) {
	RchitStage st(in);
	st.main();
}
}  // namespace glslref
