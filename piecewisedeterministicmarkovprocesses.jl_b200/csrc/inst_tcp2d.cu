// one translation unit per model so that nvcc compiles the kernel variants in parallel
#include "ensemble_kernel.cuh"
#include "models.cuh"
namespace pdmp {
ModelEntry entry_tcp2d(int id) { return make_entry<Tcp2dDev>(id); }
}  // namespace pdmp
