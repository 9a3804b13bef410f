// Reader for iharm3d HDF5 dumps (header/{n1,n2,n3,gam,metric,n_prim,prim_names,geom/...}, t, prims), producing the
// same host arrays as the .athdf reader: one block; modified Kerr-Schild coordinates x1 = ln r,
// theta = pi x2 + (1 - h)/2 sin(2 pi x2) are converted to spherical Kerr-Schild ones, normal-frame primitive
// three-vectors to the standard normal frame / coordinate-frame field.  With simulation_coord = fmks ("funky" MKS,
// theta depends on x1 and x2) the coordinates stay native and a table mapping (r, theta) back to (x1, x2) goes
// with them (reference simulation_reader.cpp:362-432,622-660,782-807,1296-1424; simulation_geometry.cpp:29-240,330-471).
#pragma once
#include <string>

#include "athdf.hpp"

namespace blh {

struct Iharm3dExpect {
  bool fmks = false;            // simulation_coord = fmks (else sks)
  double simulation_a = 0.0;
  // in: the input file's values where set; out: the values to use (the dump's header/gam, gam_p, gam_e otherwise)
  bool gamma_set = false, gamma_i_set = false, gamma_e_set = false, need_gamma_ie = false;
  double plasma_gamma = 0.0, plasma_gamma_i = 0.0, plasma_gamma_e = 0.0;
};

// kappa_name: electron-entropy variable when plasma_model = code_kappa ("" = none).  reuse_layout as for read_athdf.
void read_iharm3d(const std::string &path, const std::string &kappa_name, bool reuse_layout, Iharm3dExpect &expect,
                  AthenaGrid &grid);
// time `t` only
double read_iharm3d_time(const std::string &path);
// adiabatic indices only: fills expect.plasma_gamma[_i,_e] as read_iharm3d would, without warnings
void read_iharm3d_gammas(const std::string &path, Iharm3dExpect &expect);

}  // namespace blh
