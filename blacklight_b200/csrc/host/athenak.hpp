// Reader for AthenaK binary dumps ("Athena binary output version=1.1"): ascii pre-header, the run's input
// parameters as text, then one record per MeshBlock (cell index bounds, logical location, level, face
// positions, cell data variable by variable).  Produces the same host arrays as the .athdf reader: uniform
// Cartesian Kerr-Schild blocks with faces rebuilt from the block edges in double, cell centres as face
// averages, `eint` turned into pressure (reference simulation_reader.cpp:434-588,915-1131,1225-1290).
#pragma once
#include <string>

#include "athdf.hpp"

namespace blh {

// What the reader checks the dump's own parameters against (a mismatch is warned about and ignored) and
// the adiabatic index: in = the input file's value if gamma_set, out = the value to use.
struct AthenaKExpect {
  double simulation_a = 0.0, simulation_m_msun = 0.0, simulation_rho_cgs = 0.0, plasma_mu = 0.0;
  bool gamma_set = false;
  double plasma_gamma = 0.0;
};

// kappa_name: electron-entropy variable when plasma_model = code_kappa ("" = none).  reuse_layout as for read_athdf
// (layout, variable positions and the adiabatic index are taken from the first snapshot only).
void read_athenak(const std::string &path, const std::string &kappa_name, bool reuse_layout, AthenaKExpect &expect,
                  AthenaGrid &grid);
// time and adiabatic index (<mhd> gamma) from the header only
void read_athenak_header(const std::string &path, double *time, double *gamma_adi);

}  // namespace blh
