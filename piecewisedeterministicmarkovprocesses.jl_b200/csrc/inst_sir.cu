// one translation unit per model so that nvcc compiles the kernel variants in parallel
#include "ensemble_kernel.cuh"
#include "models.cuh"
namespace pdmp {
ModelEntry entry_sir(int id) { return make_entry<SirChvDev>(id); }
}  // namespace pdmp
