#include "lsq_launch.cuh"
namespace b200opt {
int lsq_launch_gauss2(const LsqArgs& a) { return lsq_launch_model<Gauss2>(a); }
}  // namespace b200opt
