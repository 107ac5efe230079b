// ORACLE -- TEST INFRASTRUCTURE ONLY. src/shaders/ray.rmiss (miss index 0), translated by glsl2cpp.py.
#include "stage_common.h"
namespace glslref {
struct RmissStage : Stage {
	using Stage::Stage;
#include "gen/ray.rmiss.inc"
};
void run_rmiss(const Stage::Inputs& in) {
	RmissStage st(in);
	st.main();
}
}  // namespace glslref
