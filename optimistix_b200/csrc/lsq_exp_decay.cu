#include "lsq_launch.cuh"
namespace b200opt {
int lsq_launch_exp_decay(const LsqArgs& a) { return lsq_launch_model<ExpDecay>(a); }
}  // namespace b200opt
