// Explicit instantiation of the engine for the pendulum model (one translation unit per model: parallel builds).
#include "rnmpc_engine.cuh"
namespace rnmpc {
IEngine* make_pendulum(const rnmpc_config* cfg, int cap, int device) { return make_engine<PendulumModel>(cfg, cap, device); }
}  // namespace rnmpc
