// Re-hosted `main` of the reference (src/blacklight.cpp:31-273): read input, build camera, trace,
// per snapshot {read grid, radiate level by level with adaptive refinement}, write output, report
// the same five-line timing summary.  All heavy work goes through the C ABI.
#pragma once
#include <string>

namespace blh {

struct RunTimings {
  double total = 0, geodesic = 0, read = 0, sample = 0, image = 0, render = 0;  // seconds (wall)
  double gpu_geodesic_ms = 0, gpu_radiation_ms = 0, gpu_refine_ms = 0;          // CUDA-event sums
  long long rays = 0, samples = 0;
};

// Throws blh::Error.  device < 0: use BLACKLIGHT_DEVICE or 0.
RunTimings run_input_file(const std::string &path, int device, bool quiet);

}  // namespace blh
