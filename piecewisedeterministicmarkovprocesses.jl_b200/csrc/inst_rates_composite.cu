// one translation unit per model so that nvcc compiles the kernel variants in parallel
#include "ensemble_kernel.cuh"
#include "models.cuh"
namespace pdmp {
ModelEntry entry_rates_composite(int id) { return make_entry<RatesCompositeDev>(id); }
}  // namespace pdmp
