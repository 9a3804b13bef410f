// InputFile -> (CameraSetup, bl_params): the parameter copying and validation the reference does in
// its GeodesicIntegrator and RadiationIntegrator constructors (geodesic_integrator.cpp:23-157,
// radiation_integrator.cpp:26-541), producing the POD the C ABI takes.
#pragma once
#include <string>
#include <vector>

#include "../../../include/blacklight_b200.h"
#include "camera.hpp"
#include "input_file.hpp"

namespace blh {

struct RunConfig {
  bl_params params;        // camera frame fields filled in
  CameraSetup camera;
  CameraFrame frame;
  std::vector<double> frequencies;
  // output / reader side
  int output_format = 0;   // 0 npz, 1 npy, 2 raw
  std::string output_file;
  bool output_camera = false;
  int simulation_format = 0;  // 0 athena, 1 athenak, 2 iharm3d, 3 harm3d
  std::string simulation_file, simulation_kappa_name;
  bool simulation_multiple = false;
  int simulation_start = 0, simulation_end = 0;
  bool gamma_set = false, gamma_i_set = false, gamma_e_set = false;
  bool checkpoint_geodesic_save = false, checkpoint_geodesic_load = false, checkpoint_sample_save = false;
  std::string checkpoint_geodesic_file, checkpoint_sample_file;
  // slow light (simulation_reader.cpp:64-82, output_writer.cpp:104)
  double slow_t_start = 0.0, slow_dt = 0.0;
  int slow_offset = 0;
  int num_runs = 1;
  int num_threads = 0;
};

RunConfig make_config(const InputFile &in);

// What bl_set_camera takes: the frame InitializeCamera leaves plus the per-pixel parameters (camera.cpp:53-380)
bl_camera make_bl_camera(const RunConfig &cfg);

}  // namespace blh
