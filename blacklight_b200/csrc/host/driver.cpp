#include "driver.hpp"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <thread>
#include <vector>

#include "athdf.hpp"
#include "snapshot.hpp"
#include "config.hpp"
#include "npz_writer.hpp"

namespace blh {

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct CtxGuard {
  bl_ctx *ctx = nullptr;
  ~CtxGuard() { bl_destroy(ctx); }
};

void check(bl_ctx *ctx, int rc) {
  if (rc != BL_OK) throw Error(bl_last_error(ctx));
}

struct LevelData {
  std::vector<int32_t> locs;              // (B,2) block coordinates (v,u)
  std::vector<double> pos, dir, factor;   // camera arrays
  std::vector<double> image, render;
  std::vector<uint8_t> flags;             // refinement flags (B)
  long long rays = 0;
  int blocks = 0;
};

// Binary dump in the reference's Array<> format: five int32 extents n1..n5 (fastest first) then data
// (reference utils/file_io.cpp:64-75); used by the checkpoint_*_save options.
template <typename T>
void write_array(std::ofstream &out, const T *data, const int n[5]) {
  out.write(reinterpret_cast<const char *>(n), 5 * sizeof(int));
  size_t total = (size_t)n[0] * n[1] * n[2] * n[3] * n[4];
  out.write(reinterpret_cast<const char *>(data), (std::streamsize)(total * sizeof(T)));
}

// image plane(s) of one named quantity: slots [slot0 + l*stride_l] for l < F, each `pix` long
std::vector<uint8_t> planes_npy(const std::vector<double> &image, long long pix, int F, int slot0, int slot_stride,
                                const std::vector<int> &plane_shape) {
  std::vector<double> tmp((size_t)F * pix);
  for (int l = 0; l < F; l++)
    std::memcpy(tmp.data() + (size_t)l * pix, image.data() + (size_t)(slot0 + l * slot_stride) * pix, (size_t)pix * sizeof(double));
  std::vector<int> shape;
  if (F > 1) shape.push_back(F);
  shape.insert(shape.end(), plane_shape.begin(), plane_shape.end());
  return npy_bytes(tmp.data(), shape);
}

struct Offsets {
  int time, length, lambda, emission, tau, lambda_ave, emission_ave, tau_int, crossings, total;
};

Offsets image_offsets(const bl_params &p) {
  Offsets o{};
  int q = 0, F = p.image_num_frequencies;
  bool sim = p.model_type == BL_MODEL_SIMULATION;
  if (p.image_light) q += F * (sim && p.image_polarization ? 4 : 1);
  o.time = q; if (p.image_time) q += 1;
  o.length = q; if (p.image_length) q += 1;
  o.lambda = q; if (p.image_lambda) q += F;
  o.emission = q; if (p.image_emission) q += F;
  o.tau = q; if (p.image_tau) q += F;
  o.lambda_ave = q; if (sim && p.image_lambda_ave) q += 7 * F;
  o.emission_ave = q; if (sim && p.image_emission_ave) q += 7 * F;
  o.tau_int = q; if (sim && p.image_tau_int) q += 7 * F;
  o.crossings = q; if (p.image_crossings) q += 1;
  o.total = q;
  return o;
}

void add_level_arrays(NpzWriter &npz, const RunConfig &cfg, const LevelData &L, int level) {
  const bl_params &p = cfg.params;
  const bool sim = p.model_type == BL_MODEL_SIMULATION;
  const bool pol = sim && p.image_light && p.image_polarization;
  const int F = p.image_num_frequencies, res = p.camera_resolution, bs = p.adaptive_block_size;
  const Offsets o = image_offsets(p);
  static const char *cell_names[7] = {"rho", "n_e", "p_gas", "Theta_e", "B", "sigma", "beta_inverse"};
  std::vector<int> plane = level == 0 ? std::vector<int>{res, res} : std::vector<int>{L.blocks, bs, bs};
  auto name = [&](const std::string &base) {
    return level == 0 ? base : "adaptive_" + base + "_" + std::to_string(level);
  };
  if (level > 0) npz.add("adaptive_block_locs_" + std::to_string(level), npy_bytes(L.locs.data(), {L.blocks, 2}));
  if (cfg.output_camera) {
    std::vector<int> shape = plane;
    shape.push_back(4);
    if (cfg.camera.type == 0) npz.add(name("positions"), npy_bytes(L.pos.data(), shape));
    else npz.add(name("directions"), npy_bytes(L.dir.data(), shape));
  }
  const long long pix = L.rays;
  if (p.image_light) {
    int stride = pol ? 4 : 1;
    npz.add(name("I_nu"), planes_npy(L.image, pix, F, 0, stride, plane));
    if (pol) {
      npz.add(name("Q_nu"), planes_npy(L.image, pix, F, 1, stride, plane));
      npz.add(name("U_nu"), planes_npy(L.image, pix, F, 2, stride, plane));
      npz.add(name("V_nu"), planes_npy(L.image, pix, F, 3, stride, plane));
    }
  }
  if (p.image_time) npz.add(name("time"), planes_npy(L.image, pix, 1, o.time, 1, plane));
  if (p.image_length) npz.add(name("length"), planes_npy(L.image, pix, 1, o.length, 1, plane));
  if (p.image_lambda) npz.add(name("lambda"), planes_npy(L.image, pix, F, o.lambda, 1, plane));
  if (p.image_emission) npz.add(name("emission"), planes_npy(L.image, pix, F, o.emission, 1, plane));
  if (p.image_tau) npz.add(name("tau"), planes_npy(L.image, pix, F, o.tau, 1, plane));
  if (sim && p.image_lambda_ave)
    for (int q = 0; q < 7; q++) npz.add(name(std::string("lambda_ave_") + cell_names[q]), planes_npy(L.image, pix, F, o.lambda_ave + q, 7, plane));
  if (sim && p.image_emission_ave)
    for (int q = 0; q < 7; q++) npz.add(name(std::string("emission_ave_") + cell_names[q]), planes_npy(L.image, pix, F, o.emission_ave + q, 7, plane));
  if (sim && p.image_tau_int)
    for (int q = 0; q < 7; q++) npz.add(name(std::string("tau_int_") + cell_names[q]), planes_npy(L.image, pix, F, o.tau_int + q, 7, plane));
  if (p.image_crossings) npz.add(name("crossings"), planes_npy(L.image, pix, 1, o.crossings, 1, plane));
  if (sim && p.render_num_images > 0) {
    std::vector<int> shape = {p.render_num_images, 3};
    shape.insert(shape.end(), plane.begin(), plane.end());
    npz.add(name("rendering"), npy_bytes(L.render.data(), shape));
  }
}

void write_output(const RunConfig &cfg, const std::vector<LevelData> &levels, int num_levels, int snapshot) {
  const bl_params &p = cfg.params;
  const bool sim = p.model_type == BL_MODEL_SIMULATION;
  std::string path = cfg.output_file;
  if (sim && cfg.simulation_multiple)
    path = format_numbered(cfg.output_file, snapshot + (p.slow_light_on ? cfg.slow_offset : cfg.simulation_start), "output_file");
  const LevelData &root = levels[0];
  const Offsets o = image_offsets(p);
  if (cfg.output_format == 0) {
    NpzWriter npz;
    double mass = p.mass_msun, width = p.camera_width;
    npz.add("mass_msun", npy_bytes(&mass, {1}));
    npz.add("width", npy_bytes(&width, {1}));
    npz.add("frequency", npy_bytes(cfg.frequencies.data(), {p.image_num_frequencies}));
    int32_t nl = num_levels;
    npz.add("adaptive_num_levels", npy_bytes(&nl, {1}));
    if (p.adaptive_max_level > 0) {
      std::vector<int32_t> counts;
      for (int l = 0; l <= num_levels; l++) counts.push_back(levels[(size_t)l].blocks);
      npz.add("adaptive_num_blocks", npy_bytes(counts.data(), {num_levels + 1}));
    }
    for (int l = 0; l <= num_levels; l++) add_level_arrays(npz, cfg, levels[(size_t)l], l);
    npz.write(path);
  } else {
    std::ofstream out(path, std::ios::binary);
    if (!out.is_open()) throw Error("Could not open output file.");
    if (cfg.output_format == 1) {
      std::vector<uint8_t> npy = npy_bytes(root.image.data(), {o.total, p.camera_resolution, p.camera_resolution});
      out.write(reinterpret_cast<const char *>(npy.data()), (std::streamsize)npy.size());
    } else {
      out.write(reinterpret_cast<const char *>(root.image.data()), (std::streamsize)(root.image.size() * sizeof(double)));
    }
  }
}

void validate_output_options(const RunConfig &cfg) {
  // reference output_writer.cpp:58-107
  const bl_params &p = cfg.params;
  const bool sim = p.model_type == BL_MODEL_SIMULATION;
  if (p.image_num_frequencies > 1 && cfg.output_format != 0) throw Error("Only npz support multiple frequencies.");
  if (p.image_light && sim && p.image_polarization && cfg.output_format == 2) throw Error("Only npz or npy outputs support polarization.");
  if ((p.image_time || p.image_length || p.image_lambda || p.image_emission || p.image_tau || p.image_lambda_ave ||
       p.image_emission_ave || p.image_tau_int || p.image_crossings) && cfg.output_format != 0)
    throw Error("Only npz outputs support non-light images.");
  if (p.render_num_images > 0 && cfg.output_format != 0) throw Error("Only npz outputs support rendering.");
  if (p.adaptive_max_level > 0 && cfg.output_format != 0) throw Error("Only npz outputs support adaptive ray tracing.");
}

// Level-0 geodesic checkpoint in the reference's byte format (geodesic_checkpoint.cpp:28-59), so CPU
// and GPU runs can exchange geodesics: camera frame vectors, camera arrays, frequencies, momentum
// factors, geodesic_num_steps, sample_flags/num/pos/dir/len.
void save_geodesic_checkpoint(bl_ctx *ctx, const RunConfig &cfg, const LevelData &root, int S) {
  size_t N = (size_t)root.rays, ns = N * (size_t)(S > 0 ? S : 1);
  std::vector<uint8_t> flags(N);
  std::vector<int32_t> num(N);
  std::vector<double> pos(ns * 4), dir(ns * 4), len(ns);
  check(ctx, bl_download_samples(ctx, 0, flags.data(), num.data(), pos.data(), dir.data(), len.data()));
  std::ofstream out(cfg.checkpoint_geodesic_file, std::ios::binary);
  if (!out.is_open()) throw Error("Could not open geodesic checkpoint file.");
  const CameraFrame &f = cfg.frame;
  for (const double *v : {f.x, f.u_con, f.u_cov, f.norm_con, f.norm_con_c, f.hor_con_c, f.vert_con_c})
    out.write(reinterpret_cast<const char *>(v), 4 * sizeof(double));
  int F = cfg.params.image_num_frequencies;
  int n_cam[5] = {4, (int)N, 1, 1, 1}, n_f[5] = {F, 1, 1, 1, 1}, n_ray[5] = {(int)N, 1, 1, 1, 1};
  int n_vec[5] = {4, S, (int)N, 1, 1}, n_len[5] = {S, (int)N, 1, 1, 1};
  write_array(out, root.pos.data(), n_cam);
  write_array(out, root.dir.data(), n_cam);
  write_array(out, cfg.frequencies.data(), n_f);
  write_array(out, root.factor.data(), n_ray);
  out.write(reinterpret_cast<const char *>(&S), sizeof(int));
  write_array(out, flags.data(), n_ray);
  write_array(out, num.data(), n_ray);
  write_array(out, pos.data(), n_vec);
  write_array(out, dir.data(), n_vec);
  write_array(out, len.data(), n_len);
}

// Level-0 sampling checkpoint in the reference's byte format (sample_checkpoint.cpp:22-39): sample_inds (N,S,4) int32,
// sample_fracs (N,S,3) when interpolating, sample_nan, sample_fallback (N,S) bytes -- what the fused kernel's parity
// taps recorded during the first bl_radiate_level.  Entries the reference leaves unset (cut / off-grid samples,
// n >= sample_num) hold -1 / 0 here and whatever its allocator returned there.
void save_sample_checkpoint(bl_ctx *ctx, const RunConfig &cfg, const LevelData &root, int S) {
  if (cfg.params.simulation_block_interp)
    throw Error("checkpoint_sample_save with simulation_block_interp is outside the B200 hot-path scope.");
  const bool interp = cfg.params.simulation_interp != 0;
  size_t N = (size_t)root.rays, ns = N * (size_t)(S > 0 ? S : 1);
  std::vector<int32_t> inds(ns * 4);
  std::vector<double> fracs(interp ? ns * 3 : 0);
  std::vector<uint8_t> nan_(ns), fallback(ns);
  check(ctx, bl_download_sample_inds(ctx, 0, inds.data(), interp ? fracs.data() : nullptr, nan_.data(), nullptr, fallback.data()));
  std::ofstream out(cfg.checkpoint_sample_file, std::ios::binary);
  if (!out.is_open()) throw Error("Could not open sample checkpoint file.");
  int n_inds[5] = {4, S, (int)N, 1, 1}, n_fracs[5] = {3, S, (int)N, 1, 1}, n_flag[5] = {S, (int)N, 1, 1, 1};
  write_array(out, inds.data(), n_inds);
  if (interp) write_array(out, fracs.data(), n_fracs);
  write_array(out, nan_.data(), n_flag);
  write_array(out, fallback.data(), n_flag);
}

// Inverse of save_geodesic_checkpoint: read a level-0 geodesic checkpoint written by the reference (or by us)
// and hand its samples to the device instead of tracing (geodesic_checkpoint.cpp:77-108).
template <typename T>
void read_array(std::ifstream &in, std::vector<T> &data, int n[5]) {
  in.read(reinterpret_cast<char *>(n), 5 * sizeof(int));
  if (!in) throw Error("Geodesic checkpoint file is truncated.");
  size_t total = (size_t)n[0] * n[1] * n[2] * n[3] * n[4];
  data.resize(total);
  in.read(reinterpret_cast<char *>(data.data()), (std::streamsize)(total * sizeof(T)));
  if (!in) throw Error("Geodesic checkpoint file is truncated.");
}

void load_geodesic_checkpoint(bl_ctx *ctx, const RunConfig &cfg, LevelData &root, bl_level_stats *st) {
  std::ifstream in(cfg.checkpoint_geodesic_file, std::ios::binary);
  if (!in.is_open()) throw Error("Could not open geodesic checkpoint file.");
  double frame[28];
  in.read(reinterpret_cast<char *>(frame), sizeof frame);   // cam_x ... vert_con_c: recomputed from the input file
  int n[5];
  std::vector<double> freqs, pos, dir, len;
  std::vector<uint8_t> flags;
  std::vector<int32_t> num;
  read_array(in, root.pos, n);
  long long N = n[1];
  read_array(in, root.dir, n);
  read_array(in, freqs, n);
  read_array(in, root.factor, n);
  int S = 0;
  in.read(reinterpret_cast<char *>(&S), sizeof(int));
  read_array(in, flags, n);
  read_array(in, num, n);
  read_array(in, pos, n);
  read_array(in, dir, n);
  read_array(in, len, n);
  if (N != (long long)cfg.params.camera_resolution * cfg.params.camera_resolution || (long long)root.factor.size() != N ||
      (long long)root.pos.size() != 4 * N || (long long)root.dir.size() != 4 * N)
    throw Error("Geodesic checkpoint does not match camera_resolution.");
  if (S <= 0 || (long long)flags.size() != N || (long long)num.size() != N || (long long)len.size() != N * S ||
      (long long)pos.size() != 4 * N * S || (long long)dir.size() != 4 * N * S)
    throw Error("Geodesic checkpoint does not match its own sample count.");
  root.rays = N;
  check(ctx, bl_upload_samples(ctx, 0, root.pos.data(), root.dir.data(), root.factor.data(), N, S, flags.data(), num.data(),
                               pos.data(), dir.data(), len.data(), st));
}

// ---- one bl_ctx per GPU --------------------------------------------------------------------------------------------
// The reference parallelises inside main with OpenMP threads over pixels (blacklight.cpp:77,94,204,221); here the same
// main drives one context per device from one host thread each.  Rays never interact, so a level is split into units
// -- image rows of a plain frame, refinement blocks of an adaptive one (the root level included: level0_block_major)
// -- dealt round-robin over the devices (ray cost varies strongly across the image); the grid is replicated.  What is
// exchanged: per level the refinement flags (every device flags its own blocks, the host merges them into the one
// vector the child list is derived from, camera.cpp:445-459) and, per level and snapshot, each device's image part,
// which its bl_radiate_level copies device-to-host straight into its slice of a host buffer and the device's thread
// then scatters into the frame in the reference's pixel order.  Results are bitwise independent of the device count.
struct Worker {
  int device = 0;
  bl_ctx *ctx = nullptr;
  std::vector<double> pos, dir, factor;   // camera arrays of the level being traced
  std::vector<double> image, render;      // this device's part of the level being radiated
  std::vector<int32_t> locs;              // this device's blocks of that level
  std::vector<uint8_t> flags;
  std::vector<std::vector<long long>> units;   // per level: ids of the rows / blocks this device owns
  bl_level_stats st{};
  double ms_geodesic = 0.0, ms_radiation = 0.0, ms_refine = 0.0;
  long long samples = 0;
};

struct Workers {
  std::vector<Worker> w;
  ~Workers() {
    for (Worker &x : w) bl_destroy(x.ctx);
  }
  // fn(worker) on one host thread per device; the first exception is rethrown on the caller
  template <typename F>
  void each(int host_threads, F fn) {
    if (w.size() == 1) {
      fn(w[0]);
      return;
    }
    std::vector<std::exception_ptr> err(w.size());
    std::vector<std::thread> threads;
    const int per = std::max(1, host_threads / (int)w.size());
    for (size_t i = 0; i < w.size(); i++)
      threads.emplace_back([&, i] {
        try {
          omp_set_num_threads(per);   // the camera loops inside run on this thread's own OpenMP team
          fn(w[i]);
        } catch (...) {
          err[i] = std::current_exception();
        }
      });
    for (std::thread &t : threads) t.join();
    for (std::exception_ptr &e : err)
      if (e) std::rethrow_exception(e);
  }
};

// BLACKLIGHT_DEVICES = "all" | "0-7" | "0,2,3" (a list of CUDA ordinals); unset: the single device BLACKLIGHT_DEVICE or 0
std::vector<int> devices_from_environment(int device) {
  std::vector<int> out;
  if (device >= 0) return {device};
  const char *list = std::getenv("BLACKLIGHT_DEVICES");
  if (list && *list) {
    std::string s(list);
    if (s == "all") {
      int n = bl_device_count();
      for (int d = 0; d < n; d++) out.push_back(d);
    } else {
      std::stringstream ss(s);
      std::string item;
      while (std::getline(ss, item, ',')) {
        size_t dash = item.find('-');
        try {
          if (dash == std::string::npos) {
            out.push_back(std::stoi(item));
          } else {
            int lo = std::stoi(item.substr(0, dash)), hi = std::stoi(item.substr(dash + 1));
            for (int d = lo; d <= hi; d++) out.push_back(d);
          }
        } catch (const std::exception &) {
          throw Error("Could not parse BLACKLIGHT_DEVICES.");
        }
      }
    }
    if (out.empty()) throw Error("BLACKLIGHT_DEVICES names no device.");
    return out;
  }
  const char *env = std::getenv("BLACKLIGHT_DEVICE");
  return {env ? std::atoi(env) : 0};
}

}  // namespace

RunTimings run_input_file(const std::string &path, int device, bool quiet) {
  return run_input_file(path, devices_from_environment(device), quiet);
}

RunTimings run_input_file(const std::string &path, const std::vector<int> &devices_in, bool quiet) {
  RunTimings T;
  double t_begin = now_s();
  InputFile in(path);
  RunConfig cfg = make_config(in);
  validate_output_options(cfg);
  bl_params &p = cfg.params;
  const bool sim = p.model_type == BL_MODEL_SIMULATION;
  std::unique_ptr<SnapshotReader> reader;
  if (sim) {
    reader.reset(new SnapshotReader(cfg));
    p.plasma_gamma = reader->plasma_gamma();
    p.plasma_gamma_i = reader->plasma_gamma_i();
    p.plasma_gamma_e = reader->plasma_gamma_e();
  }
  auto read_snapshot = [&](const std::string &file, bool reuse, AthenaGrid &into) { reader->read(file, reuse, into); };
  auto snapshot_time_of = [&](const std::string &file) { return reader->time_of(file); };
  std::vector<int> devices = devices_in;
  if (devices.empty()) throw Error("No device given.");
  // checkpoints hold whole levels in the reference's layouts: they go through a single device
  if (devices.size() > 1 && (cfg.checkpoint_geodesic_save || cfg.checkpoint_geodesic_load || cfg.checkpoint_sample_save)) {
    warning("Checkpoints are exchanged through a single device; ignoring all but the first of BLACKLIGHT_DEVICES.");
    devices.resize(1);
  }
  const int N = (int)devices.size();
  const int host_threads = cfg.num_threads > 0 ? cfg.num_threads : omp_get_max_threads();
  const bool adaptive = p.adaptive_max_level > 0;
  const int bs = p.adaptive_block_size, res = p.camera_resolution;
  if (N > 1 && adaptive) p.level0_block_major = 1;   // root blocks are handed over block by block, like refined ones

  // camera pixels are generated on the device unless BLACKLIGHT_HOST_CAMERA=1 (A/B and parity checks)
  const char *host_camera_env = std::getenv("BLACKLIGHT_HOST_CAMERA");
  const bool host_camera = host_camera_env && std::atoi(host_camera_env) != 0;
  const bl_camera cam = make_bl_camera(cfg);

  Workers W;
  W.w.resize((size_t)N);
  for (int i = 0; i < N; i++) {
    W.w[(size_t)i].device = devices[(size_t)i];
    W.w[(size_t)i].units.resize((size_t)p.adaptive_max_level + 1);
  }
  W.each(host_threads, [&](Worker &w) {
    bl_params q = p;
    q.device = w.device;
    if (bl_create(&q, &w.ctx) != BL_OK) throw Error(bl_last_error(nullptr));
    check(w.ctx, bl_set_camera(w.ctx, &cam));
  });
  bl_ctx *ctx0 = W.w[0].ctx;
  const int Q = bl_image_num_quantities(ctx0);
  const int R = sim ? p.render_num_images : 0;
  std::vector<LevelData> levels((size_t)p.adaptive_max_level + 1);
  const bool block_units = N > 1 && adaptive;                  // units of level 0: blocks, else rows
  const long long bs2 = (long long)bs * bs;
  auto unit_rays = [&](int level) { return level == 0 && !block_units ? (long long)res : bs2; };

  // units of a level with `count` of them, dealt round-robin
  auto deal = [&](int level, long long count) {
    for (int i = 0; i < N; i++) {
      std::vector<long long> &u = W.w[(size_t)i].units[(size_t)level];
      u.clear();
      for (long long k = i; k < count; k += N) u.push_back(k);
    }
  };
  // trace this device's share of a level: the pixels of its rows / blocks are generated on the device
  // (bl_trace_level_pixels; BLACKLIGHT_HOST_CAMERA=1 builds them here and uploads them, as round 1 did)
  auto trace_share = [&](Worker &w, int level, const LevelData &L) {
    const std::vector<long long> &u = w.units[(size_t)level];
    const long long rays = (long long)u.size() * unit_rays(level);
    const bool rows = level == 0 && !block_units;
    w.locs.resize(rows ? u.size() : u.size() * 2);
    for (size_t k = 0; k < u.size(); k++) {
      if (rows) {
        w.locs[k] = (int32_t)u[k];
      } else {
        w.locs[2 * k] = L.locs[2 * (size_t)u[k]];
        w.locs[2 * k + 1] = L.locs[2 * (size_t)u[k] + 1];
      }
    }
    if (!host_camera) {
      check(w.ctx, bl_trace_level_pixels(w.ctx, level, rows ? BL_PIXELS_ROWS : BL_PIXELS_BLOCKS, w.locs.data(), (int64_t)u.size(), &w.st));
    } else {
      w.pos.resize((size_t)rays * 4);
      w.dir.resize((size_t)rays * 4);
      w.factor.resize((size_t)rays);
      if (rows)
        camera_rows(cfg.camera, cfg.frame, u.data(), (long long)u.size(), w.pos.data(), w.dir.data(), w.factor.data());
      else
        camera_blocks(cfg.camera, cfg.frame, level, bs, w.locs.data(), (long long)u.size(), w.pos.data(), w.dir.data(), w.factor.data());
      check(w.ctx, bl_trace_level(w.ctx, level, w.pos.data(), w.dir.data(), w.factor.data(), rays, &w.st));
    }
    w.ms_geodesic += w.st.ms_geodesic;
  };
  auto bad_geodesics_warning = [&](long long rays) {
    long long bad = 0;
    for (Worker &w : W.w) bad += w.st.num_bad_geodesics;
    if (bad > 0) warning(std::to_string(bad) + " out of " + std::to_string(rays) + " geodesics terminate unexpectedly.");
  };

  // level 0 camera + geodesics (GeodesicIntegrator::Integrate)
  double t0 = now_s();
  LevelData &root = levels[0];
  root.rays = (long long)res * res;
  if (adaptive) {
    int nb = res / bs;
    root.blocks = nb * nb;
    root.locs.resize((size_t)root.blocks * 2);
    for (int v = 0, b = 0; v < nb; v++)
      for (int u = 0; u < nb; u++, b++) {
        root.locs[2 * (size_t)b] = v;
        root.locs[2 * (size_t)b + 1] = u;
      }
  }
  bl_level_stats st{};
  int level0_steps = 0;
  if (N == 1) {
    // single device: the whole raster, exactly as the reference's main
    // (the host copy of the camera arrays is only needed where it is written out)
    if (!cfg.checkpoint_geodesic_load && (host_camera || cfg.output_camera || cfg.checkpoint_geodesic_save))
      camera_root(cfg.camera, cfg.frame, root.pos, root.dir, root.factor);
    if (cfg.checkpoint_geodesic_load)
      load_geodesic_checkpoint(ctx0, cfg, root, &st);
    else if (host_camera)
      check(ctx0, bl_trace_level(ctx0, 0, root.pos.data(), root.dir.data(), root.factor.data(), root.rays, &st));
    else
      check(ctx0, bl_trace_level_pixels(ctx0, 0, BL_PIXELS_ROWS, nullptr, res, &st));
    W.w[0].st = st;
    W.w[0].ms_geodesic += st.ms_geodesic;
    bad_geodesics_warning(root.rays);
    if (cfg.checkpoint_geodesic_save) save_geodesic_checkpoint(ctx0, cfg, root, st.geodesic_num_steps);
    level0_steps = st.geodesic_num_steps;
  } else {
    deal(0, block_units ? root.blocks : res);
    W.each(host_threads, [&](Worker &w) { trace_share(w, 0, root); });
    bad_geodesics_warning(root.rays);
    if (cfg.output_camera) camera_root(cfg.camera, cfg.frame, root.pos, root.dir, root.factor);
  }
  T.geodesic += now_s() - t0;

  AthenaGrid grid;
  // slow light: the reader's sliding window of snapshots (simulation_reader.cpp:211-303), kept resident in HBM.
  // Window entry t (0 = latest) lives in device slot window_slot[t]; shifting the window permutes the slots.
  std::vector<int32_t> window_slot;
  std::vector<double> window_time;
  int latest_file_number = -1;
  bool first_read = true;
  for (int n = 0; n < cfg.num_runs; n++) {
    if (sim && p.slow_light_on) {
      t0 = now_s();
      const int chunk = p.slow_chunk_size;
      const double tol = p.extrapolation_tolerance;
      const double snapshot_time = cfg.slow_t_start + cfg.slow_dt * n;
      double latest_time = first_read ? snapshot_time - 2.0 * tol : window_time[0];
      int latest_old = -1;
      if (first_read) {
        latest_file_number = cfg.simulation_start + chunk - 2;
        window_slot.resize((size_t)chunk);
        window_time.assign((size_t)chunk, 0.0);
        for (int t = 0; t < chunk; t++) window_slot[(size_t)t] = t;
      } else {
        latest_old = latest_file_number;
      }
      while (latest_time < snapshot_time && latest_file_number < cfg.simulation_end) {
        latest_file_number++;
        latest_time = snapshot_time_of(format_numbered(cfg.simulation_file, latest_file_number, "simulation_file"));
      }
      if (latest_time < snapshot_time - tol) {
        std::ostringstream msg;
        msg << "Snapshot " << n << " at time " << snapshot_time << " would require significant extrapolation beyond file "
            << cfg.simulation_end << ".";
        throw Error(msg.str());
      } else if (latest_time < snapshot_time) {
        std::ostringstream msg;
        msg << "Snapshot " << n << " at time " << snapshot_time << " requires moderate extrapolation.";
        warning(msg.str());
      }
      int num_read;
      if (latest_file_number == latest_old) {
        num_read = 0;
      } else if (latest_file_number - chunk + 1 <= latest_old) {
        num_read = latest_file_number - latest_old;
        // entries move back by num_read; the slots of the entries that fall off the end are reused for the new ones
        std::vector<int32_t> freed(window_slot.end() - num_read, window_slot.end());
        for (int t = chunk - 1; t >= num_read; t--) {
          window_slot[(size_t)t] = window_slot[(size_t)(t - num_read)];
          window_time[(size_t)t] = window_time[(size_t)(t - num_read)];
        }
        for (int t = 0; t < num_read; t++) window_slot[(size_t)t] = freed[(size_t)t];
      } else {
        num_read = chunk;
      }
      for (int t = 0; t < num_read; t++) {
        std::string file = format_numbered(cfg.simulation_file, latest_file_number - t, "simulation_file");
        read_snapshot(file, !first_read, grid);
        first_read = false;
        window_time[(size_t)t] = grid.time;
        bl_grid_view view = grid.view();
        W.each(host_threads, [&](Worker &w) { check(w.ctx, bl_upload_grid_slice(w.ctx, &view, window_slot[(size_t)t])); });
      }
      W.each(host_threads, [&](Worker &w) {
        check(w.ctx, bl_set_time_window(w.ctx, chunk, window_slot.data(), window_time.data(), snapshot_time));
      });
      T.read += now_s() - t0;
    } else if (sim) {
      t0 = now_s();
      std::string file = cfg.simulation_file;
      if (cfg.simulation_multiple) file = format_numbered(cfg.simulation_file, cfg.simulation_start + n, "simulation_file");
      read_snapshot(file, n > 0, grid);
      bl_grid_view view = grid.view();
      W.each(host_threads, [&](Worker &w) { check(w.ctx, bl_upload_grid(w.ctx, &view)); });   // replicated on every device
      T.read += now_s() - t0;
    }
    int level = 0, num_levels = 0;
    for (;;) {
      LevelData &L = levels[(size_t)level];
      t0 = now_s();
      L.image.resize((size_t)Q * L.rays);
      if (R > 0) L.render.resize((size_t)R * 3 * L.rays);
      const bool save_sampling = sim && cfg.checkpoint_sample_save && n == 0 && level == 0;
      if (N == 1) {
        if (save_sampling) check(ctx0, bl_set_taps(ctx0, 1));
        check(ctx0, bl_radiate_level(ctx0, level, n, L.image.data(), R > 0 ? L.render.data() : nullptr, &st));
        W.w[0].st = st;
        W.w[0].ms_radiation += st.ms_radiation;
        if (level == 0 && n == 0 && st.ms_geodesic > 0 && W.w[0].ms_geodesic == 0) W.w[0].ms_geodesic += st.ms_geodesic;
        W.w[0].samples += st.num_samples;
        if (save_sampling) {   // as the reference, after the first sampling pass of level 0 (radiation_integrator.cpp:697-704)
          save_sample_checkpoint(ctx0, cfg, L, level0_steps);
          check(ctx0, bl_set_taps(ctx0, 0));
        }
      } else {
        // every device radiates its units; its thread then scatters them into the frame (disjoint destinations)
        const long long ur = unit_rays(level);
        const long long units_total = L.rays / ur;
        W.each(host_threads, [&](Worker &w) {
          const std::vector<long long> &u = w.units[(size_t)level];
          const long long rays = (long long)u.size() * ur;
          w.image.resize((size_t)Q * rays);
          if (R > 0) w.render.resize((size_t)R * 3 * rays);
          const bool first_wave_trace = w.ms_geodesic == 0;
          check(w.ctx, bl_radiate_level(w.ctx, level, n, w.image.data(), R > 0 ? w.render.data() : nullptr, &w.st));
          w.ms_radiation += w.st.ms_radiation;
          if (level == 0 && n == 0 && first_wave_trace && w.st.ms_geodesic > 0) w.ms_geodesic += w.st.ms_geodesic;
          w.samples += w.st.num_samples;
          for (int q = 0; q < Q; q++)
            for (size_t k = 0; k < u.size(); k++)
              std::memcpy(L.image.data() + ((size_t)q * units_total + (size_t)u[k]) * ur, w.image.data() + ((size_t)q * u.size() + k) * ur,
                          (size_t)ur * sizeof(double));
          for (int q = 0; q < 3 * R; q++)
            for (size_t k = 0; k < u.size(); k++)
              std::memcpy(L.render.data() + ((size_t)q * units_total + (size_t)u[k]) * ur, w.render.data() + ((size_t)q * u.size() + k) * ur,
                          (size_t)ur * sizeof(double));
        });
        if (level == 0 && block_units) {
          // the root level was radiated block by block: back to the reference's raster order (m = row * res + col)
          std::vector<double> raster(L.image.size());
          const int nb = res / bs;
#pragma omp parallel for schedule(static) collapse(2)
          for (int q = 0; q < Q; q++)
            for (int b = 0; b < nb * nb; b++) {
              const int v = b / nb, u = b % nb;
              for (int i = 0; i < bs; i++)
                std::memcpy(raster.data() + (size_t)q * L.rays + ((size_t)v * bs + i) * res + (size_t)u * bs,
                            L.image.data() + ((size_t)q * nb * nb + b) * bs2 + (size_t)i * bs, (size_t)bs * sizeof(double));
            }
          L.image.swap(raster);
          if (R > 0) {
            std::vector<double> rr(L.render.size());
            for (int q = 0; q < 3 * R; q++)
              for (int b = 0; b < nb * nb; b++) {
                const int v = b / nb, u = b % nb;
                for (int i = 0; i < bs; i++)
                  std::memcpy(rr.data() + (size_t)q * L.rays + ((size_t)v * bs + i) * res + (size_t)u * bs,
                              L.render.data() + ((size_t)q * nb * nb + b) * bs2 + (size_t)i * bs, (size_t)bs * sizeof(double));
              }
            L.render.swap(rr);
          }
        }
      }
      if (sim && p.slow_light_on) {
        // same errors / warnings as the reference's sampling stage (simulation_sampling.cpp:577-617)
        bl_slow_stats ss{};
        for (Worker &w : W.w) {
          bl_slow_stats part{};
          check(w.ctx, bl_slow_light_stats(w.ctx, level, &part));
          for (int side = 0; side < 2; side++) {
            ss.num_small[side] += part.num_small[side];
            ss.num_large[side] += part.num_large[side];
            ss.val_small[side] = std::max(ss.val_small[side], part.val_small[side]);
            ss.val_large[side] = std::max(ss.val_large[side], part.val_large[side]);
          }
        }
        const double snapshot_time = cfg.slow_t_start + cfg.slow_dt * n;
        const char *direction[2] = {"forward", "backward"};
        for (int side = 0; side < 2; side++)
          if (ss.num_large[side] > 0) {
            std::ostringstream msg;
            msg << "Snapshot " << n << " at time " << snapshot_time << " requires significant extrapolation " << direction[side]
                << " in time (" << ss.num_large[side] << "/" << L.rays << " pixels, by up to " << ss.val_large[side]
                << " gravitational times).";
            throw Error(msg.str());
          }
        for (int side = 0; side < 2; side++)
          if (ss.num_small[side] > 0) {
            std::ostringstream msg;
            msg << "Snapshot " << n << " at time " << snapshot_time << " requires moderate extrapolation " << direction[side]
                << " in time (" << ss.num_small[side] << "/" << L.rays << " pixels, by up to " << ss.val_small[side]
                << " gravitational times).";
            warning(msg.str());
          }
      }
      T.rays += L.rays;
      bool complete = true;
      if (adaptive && level < p.adaptive_max_level) {
        L.flags.assign((size_t)L.blocks, 0);
        int64_t refined = 0;
        if (N == 1) {
          check(ctx0, bl_refine_level(ctx0, level, L.locs.data(), L.blocks, L.flags.data(), &refined));
          W.w[0].ms_refine += 0.0;
        } else {
          // each device flags its own blocks; the merged vector is what every child list is derived from
          W.each(host_threads, [&](Worker &w) {
            const std::vector<long long> &u = w.units[(size_t)level];
            w.locs.resize(u.size() * 2);
            for (size_t k = 0; k < u.size(); k++) {
              w.locs[2 * k] = L.locs[2 * (size_t)u[k]];
              w.locs[2 * k + 1] = L.locs[2 * (size_t)u[k] + 1];
            }
            w.flags.assign(u.size(), 0);
            int64_t mine = 0;
            check(w.ctx, bl_refine_level(w.ctx, level, w.locs.data(), (int64_t)u.size(), w.flags.data(), &mine));
            for (size_t k = 0; k < u.size(); k++) L.flags[(size_t)u[k]] = w.flags[k];
          });
          for (uint8_t f : L.flags) refined += f ? 1 : 0;
        }
        complete = refined == 0;
      }
      T.image += now_s() - t0;
      if (complete) {
        num_levels = level;
        break;
      }
      // next level: augment camera, trace (GeodesicIntegrator::AddGeodesics)
      t0 = now_s();
      LevelData &C = levels[(size_t)level + 1];
      if (N == 1) {
        if (host_camera || cfg.output_camera) {
          camera_refined(cfg.camera, cfg.frame, level + 1, bs, L.locs, L.flags, C.locs, C.pos, C.dir, C.factor);
        } else {
          child_blocks(L.locs, L.flags, C.locs);
        }
        C.blocks = (int)(C.locs.size() / 2);
        C.rays = (long long)C.blocks * bs * bs;
        if (host_camera)
          check(ctx0, bl_trace_level(ctx0, level + 1, C.pos.data(), C.dir.data(), C.factor.data(), C.rays, &st));
        else
          check(ctx0, bl_trace_level_pixels(ctx0, level + 1, BL_PIXELS_BLOCKS, C.locs.data(), C.blocks, &st));
        W.w[0].st = st;
        W.w[0].ms_geodesic += st.ms_geodesic;
      } else {
        // child list: parents in index order x 4 children (camera.cpp:445-459); pixels are built per device
        child_blocks(L.locs, L.flags, C.locs);
        C.blocks = (int)(C.locs.size() / 2);
        C.rays = (long long)C.blocks * bs * bs;
        deal(level + 1, C.blocks);
        W.each(host_threads, [&](Worker &w) { trace_share(w, level + 1, C); });
        if (cfg.output_camera) {
          C.pos.resize((size_t)C.rays * 4);
          C.dir.resize((size_t)C.rays * 4);
          C.factor.resize((size_t)C.rays);
          camera_blocks(cfg.camera, cfg.frame, level + 1, bs, C.locs.data(), C.blocks, C.pos.data(), C.dir.data(), C.factor.data());
        }
      }
      bad_geodesics_warning(C.rays);
      T.geodesic += now_s() - t0;
      level++;
    }
    write_output(cfg, levels, num_levels, n);
  }
  for (Worker &w : W.w) {
    // device times: the slowest device bounds the run
    T.gpu_geodesic_ms = std::max(T.gpu_geodesic_ms, w.ms_geodesic);
    T.gpu_radiation_ms = std::max(T.gpu_radiation_ms, w.ms_radiation);
    T.samples += w.samples;
  }
  T.total = now_s() - t_begin;
  if (!quiet) {
    // same report as the reference (blacklight.cpp:259-269); sampling is fused into the image stage here
    std::printf("\nCalculation completed.");
    std::printf("\nElapsed time:            %.7g s", T.total);
    std::printf("\n  Integrating geodesics: %.7g s", T.geodesic);
    std::printf("\n  Reading simulation:    %.7g s", T.read);
    std::printf("\n  Sampling simulation:   %.7g s", T.sample);
    std::printf("\n  Integrating image:     %.7g s", T.image);
    std::printf("\n  Rendering:             %.7g s", T.render);
    std::printf("\n\n");
    std::printf("[B200] %d device(s): geodesic kernels %.3f ms, radiation kernels %.3f ms (slowest device), %lld rays, %lld samples\n",
                N, T.gpu_geodesic_ms, T.gpu_radiation_ms, T.rays, T.samples);
  }
  return T;
}

}  // namespace blh
