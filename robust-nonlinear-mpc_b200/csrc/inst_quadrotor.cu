// Explicit instantiation of the engine for the quadrotor model (one translation unit per model: parallel builds).
#include "rnmpc_engine.cuh"
namespace rnmpc {
IEngine* make_quadrotor(const rnmpc_config* cfg, int cap, int device) { return make_engine<QuadrotorModel>(cfg, cap, device); }
}  // namespace rnmpc
