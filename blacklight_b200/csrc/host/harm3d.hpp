// Reader for the 'harm3d' snapshot format (ascii header + float32 cell records in modified Kerr-Schild
// coordinates x1 = ln r, x2 with theta = pi x2 + (1 - h)/2 sin(2 pi x2), x3 = phi), producing the same host
// arrays as the .athdf reader: one block, coordinates and primitives converted to spherical Kerr-Schild
// (reference simulation_reader.cpp:265,360,661-720,808-848; simulation_geometry.cpp:29-92,242-327).
#pragma once
#include <string>

#include "athdf.hpp"

namespace blh {

// plasma_gamma: in = the input file's value if gamma_set, out = the value to use (the file's if !gamma_set).
// simulation_a: the input file's spin (a mismatch with the file's is warned about and ignored).
// want_kappa: plasma_model = code_kappa (a 17th float per cell).  reuse_layout as for read_athdf.
void read_harm3d(const std::string &path, bool want_kappa, bool gamma_set, double *plasma_gamma, double simulation_a,
                 bool reuse_layout, AthenaGrid &grid);
// time and adiabatic index from the header only
void read_harm3d_header(const std::string &path, double *time, double *gamma_adi);

}  // namespace blh
