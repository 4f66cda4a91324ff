// Explicit instantiation of the engine for the rocket model (one translation unit per model: parallel builds).
#include "rnmpc_engine.cuh"
namespace rnmpc {
IEngine* make_rocket(const rnmpc_config* cfg, int cap, int device) { return make_engine<RocketModel>(cfg, cap, device); }
}  // namespace rnmpc
