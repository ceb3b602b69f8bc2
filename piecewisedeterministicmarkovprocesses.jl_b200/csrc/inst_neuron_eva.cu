// one translation unit per model so that nvcc compiles the kernel variants in parallel
#include "ensemble_kernel.cuh"
#include "models.cuh"
namespace pdmp {
ModelEntry entry_neuron_eva(int id) { return make_entry<NeuronEvaDev>(id); }
}  // namespace pdmp
