// Re-hosted `main` of the reference (src/blacklight.cpp:31-273): read input, build camera, trace,
// per snapshot {read grid, radiate level by level with adaptive refinement}, write output, report
// the same five-line timing summary.  All heavy work goes through the C ABI.
#pragma once
#include <string>
#include <vector>

namespace blh {

struct RunTimings {
  double total = 0, geodesic = 0, read = 0, sample = 0, image = 0, render = 0;  // seconds (wall)
  double gpu_geodesic_ms = 0, gpu_radiation_ms = 0, gpu_refine_ms = 0;          // CUDA-event sums
  long long rays = 0, samples = 0;
};

// Throws blh::Error.  device < 0: the devices BLACKLIGHT_DEVICES names ("all", "0-7", "0,2,3"), else BLACKLIGHT_DEVICE, else 0.
RunTimings run_input_file(const std::string &path, int device, bool quiet);
// The same on an explicit list of CUDA devices: one context and one host thread per device, image rows (adaptive runs:
// refinement blocks) dealt round-robin, the grid replicated; outputs are bitwise independent of the list.
RunTimings run_input_file(const std::string &path, const std::vector<int> &devices, bool quiet);

}  // namespace blh
