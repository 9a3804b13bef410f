// Polarized (Stokes IQUV) transfer kernel -- placeholder until the coherency-tensor transport lands.
#include "rad_types.cuh"
extern "C" cudaError_t bl_launch_radiate_polarized(const RadArgs *args, int num_freq, cudaStream_t stream) {
  (void)args; (void)num_freq; (void)stream;
  return cudaErrorNotSupported;
}
