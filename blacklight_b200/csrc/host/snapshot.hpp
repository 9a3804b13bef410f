// Format dispatch of the snapshot readers -- the upload side of SimulationReader::Read
// (reference simulation_reader.cpp:200-861): Athena++ .athdf, AthenaK binary, harm3d.
#pragma once
#include <string>

#include "athdf.hpp"
#include "athenak.hpp"
#include "config.hpp"
#include "iharm3d.hpp"

namespace blh {

class SnapshotReader {
 public:
  // Reads the first snapshot's header where the format keeps the adiabatic index in the file and the input file
  // does not override it: plasma_gamma() is a kernel parameter and must be known before bl_create.
  explicit SnapshotReader(const RunConfig &cfg);
  double plasma_gamma() const { return gamma_; }
  // ion / electron indices (plasma_use_p = false): the iharm3d dump's where the input file gives none
  double plasma_gamma_i() const { return iharm_.plasma_gamma_i; }
  double plasma_gamma_e() const { return iharm_.plasma_gamma_e; }
  // reuse_layout: `grid` already holds the first snapshot's coordinates; only refresh the cell data.
  void read(const std::string &file, bool reuse_layout, AthenaGrid &grid);
  double time_of(const std::string &file) const;
  std::string first_file() const;

 private:
  const RunConfig &cfg_;
  double gamma_;
  AthenaKExpect athenak_;
  Iharm3dExpect iharm_;
};

}  // namespace blh
