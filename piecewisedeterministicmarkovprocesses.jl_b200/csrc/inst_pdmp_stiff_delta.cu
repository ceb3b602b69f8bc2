// one translation unit per model so that nvcc compiles the kernel variants in parallel
#include "ensemble_kernel.cuh"
#include "models.cuh"
namespace pdmp {
ModelEntry entry_pdmp_stiff_delta(int id) { return make_entry<StiffDeltaDev>(id); }
}  // namespace pdmp
