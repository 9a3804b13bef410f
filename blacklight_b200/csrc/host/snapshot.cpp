#include "snapshot.hpp"

#include "harm3d.hpp"
#include "input_file.hpp"

namespace blh {

namespace {
enum { kAthena = 0, kAthenaK = 1, kIharm3d = 2, kHarm3d = 3 };
}

SnapshotReader::SnapshotReader(const RunConfig &cfg) : cfg_(cfg), gamma_(cfg.params.plasma_gamma) {
  const bl_params &p = cfg.params;
  iharm_.fmks = p.simulation_coord == BL_COORD_FMKS;
  iharm_.simulation_a = p.bh_a;
  iharm_.gamma_set = cfg.gamma_set;
  iharm_.gamma_i_set = cfg.gamma_i_set;
  iharm_.gamma_e_set = cfg.gamma_e_set;
  iharm_.need_gamma_ie = p.plasma_model == BL_PLASMA_TI_TE_BETA && !p.plasma_use_p;
  iharm_.plasma_gamma = p.plasma_gamma;
  iharm_.plasma_gamma_i = p.plasma_gamma_i;
  iharm_.plasma_gamma_e = p.plasma_gamma_e;
  if (cfg.simulation_format == kIharm3d) {
    if (p.simulation_coord == BL_COORD_CKS) throw Error("Invalid simulation_coord for Harm format.");
    read_iharm3d_gammas(first_file(), iharm_);
    gamma_ = iharm_.plasma_gamma;
  } else if (p.simulation_coord == BL_COORD_FMKS) {
    throw Error("simulation_coord = fmks needs simulation_format = iharm3d.");
  }
  if (!cfg.gamma_set) {
    if (cfg.simulation_format == kHarm3d) read_harm3d_header(first_file(), nullptr, &gamma_);
    if (cfg.simulation_format == kAthenaK) read_athenak_header(first_file(), nullptr, &gamma_);
  }
  athenak_.simulation_a = cfg.params.bh_a;
  athenak_.simulation_m_msun = cfg.params.mass_msun;
  athenak_.simulation_rho_cgs = cfg.params.simulation_rho_cgs;
  athenak_.plasma_mu = cfg.params.plasma_mu;
  athenak_.gamma_set = cfg.gamma_set;
  athenak_.plasma_gamma = gamma_;
}

std::string SnapshotReader::first_file() const {
  return cfg_.simulation_multiple ? format_numbered(cfg_.simulation_file, cfg_.simulation_start, "simulation_file")
                                  : cfg_.simulation_file;
}

void SnapshotReader::read(const std::string &file, bool reuse_layout, AthenaGrid &grid) {
  const bool code_kappa = cfg_.params.plasma_model == BL_PLASMA_CODE_KAPPA;
  const std::string kappa_name = code_kappa ? cfg_.simulation_kappa_name : "";
  if (cfg_.simulation_format == kHarm3d) {
    // the header's index was taken (or compared with the input file's) above
    double g = gamma_;
    read_harm3d(file, code_kappa, cfg_.gamma_set, &g, cfg_.params.bh_a, reuse_layout, grid);
  } else if (cfg_.simulation_format == kAthenaK) {
    read_athenak(file, kappa_name, reuse_layout, athenak_, grid);
  } else if (cfg_.simulation_format == kIharm3d) {
    read_iharm3d(file, kappa_name, reuse_layout, iharm_, grid);
  } else {
    read_athdf(file, kappa_name, reuse_layout, grid);
  }
}

double SnapshotReader::time_of(const std::string &file) const {
  double t = 0.0;
  if (cfg_.simulation_format == kHarm3d) read_harm3d_header(file, &t, nullptr);
  else if (cfg_.simulation_format == kAthenaK) read_athenak_header(file, &t, nullptr);
  else if (cfg_.simulation_format == kIharm3d) t = read_iharm3d_time(file);
  else t = read_athdf_time(file);
  return t;
}

}  // namespace blh
