#include "driver.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
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
      (long long)len.size() != N * S)
    throw Error("Geodesic checkpoint does not match camera_resolution.");
  root.rays = N;
  check(ctx, bl_upload_samples(ctx, 0, root.pos.data(), root.dir.data(), root.factor.data(), N, S, flags.data(), num.data(),
                               pos.data(), dir.data(), len.data(), st));
}

}  // namespace

RunTimings run_input_file(const std::string &path, int device, bool quiet) {
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
  if (device < 0) {
    const char *env = std::getenv("BLACKLIGHT_DEVICE");
    device = env ? std::atoi(env) : 0;
  }
  p.device = device;

  CtxGuard guard;
  if (bl_create(&p, &guard.ctx) != BL_OK) throw Error(bl_last_error(nullptr));
  bl_ctx *ctx = guard.ctx;
  const int Q = bl_image_num_quantities(ctx);
  const int R = sim ? p.render_num_images : 0;
  const int bs = p.adaptive_block_size;
  std::vector<LevelData> levels((size_t)p.adaptive_max_level + 1);

  // level 0 camera + geodesics (GeodesicIntegrator::Integrate)
  double t0 = now_s();
  LevelData &root = levels[0];
  if (!cfg.checkpoint_geodesic_load) camera_root(cfg.camera, cfg.frame, root.pos, root.dir, root.factor);
  root.rays = (long long)p.camera_resolution * p.camera_resolution;
  if (p.adaptive_max_level > 0) {
    int nb = p.camera_resolution / bs;
    root.blocks = nb * nb;
    root.locs.resize((size_t)root.blocks * 2);
    for (int v = 0, b = 0; v < nb; v++)
      for (int u = 0; u < nb; u++, b++) {
        root.locs[2 * (size_t)b] = v;
        root.locs[2 * (size_t)b + 1] = u;
      }
  }
  bl_level_stats st{};
  if (cfg.checkpoint_geodesic_load)
    load_geodesic_checkpoint(ctx, cfg, root, &st);
  else
    check(ctx, bl_trace_level(ctx, 0, root.pos.data(), root.dir.data(), root.factor.data(), root.rays, &st));
  T.gpu_geodesic_ms += st.ms_geodesic;
  if (st.num_bad_geodesics > 0)
    warning(std::to_string(st.num_bad_geodesics) + " out of " + std::to_string(root.rays) + " geodesics terminate unexpectedly.");
  if (cfg.checkpoint_geodesic_save) save_geodesic_checkpoint(ctx, cfg, root, st.geodesic_num_steps);
  const int level0_steps = st.geodesic_num_steps;
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
        check(ctx, bl_upload_grid_slice(ctx, &view, window_slot[(size_t)t]));
      }
      check(ctx, bl_set_time_window(ctx, chunk, window_slot.data(), window_time.data(), snapshot_time));
      T.read += now_s() - t0;
    } else if (sim) {
      t0 = now_s();
      std::string file = cfg.simulation_file;
      if (cfg.simulation_multiple) file = format_numbered(cfg.simulation_file, cfg.simulation_start + n, "simulation_file");
      read_snapshot(file, n > 0, grid);
      bl_grid_view view = grid.view();
      check(ctx, bl_upload_grid(ctx, &view));
      T.read += now_s() - t0;
    }
    int level = 0, num_levels = 0;
    for (;;) {
      LevelData &L = levels[(size_t)level];
      t0 = now_s();
      L.image.resize((size_t)Q * L.rays);
      if (R > 0) L.render.resize((size_t)R * 3 * L.rays);
      const bool save_sampling = sim && cfg.checkpoint_sample_save && n == 0 && level == 0;
      if (save_sampling) check(ctx, bl_set_taps(ctx, 1));
      check(ctx, bl_radiate_level(ctx, level, n, L.image.data(), R > 0 ? L.render.data() : nullptr, &st));
      T.gpu_radiation_ms += st.ms_radiation;
      if (save_sampling) {   // as the reference, after the first sampling pass of level 0 (radiation_integrator.cpp:697-704)
        save_sample_checkpoint(ctx, cfg, L, level0_steps);
        check(ctx, bl_set_taps(ctx, 0));
      }
      if (sim && p.slow_light_on) {
        // same errors / warnings as the reference's sampling stage (simulation_sampling.cpp:577-617)
        bl_slow_stats ss{};
        check(ctx, bl_slow_light_stats(ctx, level, &ss));
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
      if (level == 0 && n == 0 && st.ms_geodesic > 0 && T.gpu_geodesic_ms == 0) T.gpu_geodesic_ms += st.ms_geodesic;
      T.rays += L.rays;
      T.samples += st.num_samples;
      bool complete = true;
      if (p.adaptive_max_level > 0 && level < p.adaptive_max_level) {
        L.flags.assign((size_t)L.blocks, 0);
        int64_t refined = 0;
        check(ctx, bl_refine_level(ctx, level, L.locs.data(), L.blocks, L.flags.data(), &refined));
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
      camera_refined(cfg.camera, cfg.frame, level + 1, bs, L.locs, L.flags, C.locs, C.pos, C.dir, C.factor);
      C.blocks = (int)(C.locs.size() / 2);
      C.rays = (long long)C.blocks * bs * bs;
      check(ctx, bl_trace_level(ctx, level + 1, C.pos.data(), C.dir.data(), C.factor.data(), C.rays, &st));
      T.gpu_geodesic_ms += st.ms_geodesic;
      if (st.num_bad_geodesics > 0)
        warning(std::to_string(st.num_bad_geodesics) + " out of " + std::to_string(C.rays) + " geodesics terminate unexpectedly.");
      T.geodesic += now_s() - t0;
      level++;
    }
    write_output(cfg, levels, num_levels, n);
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
    std::printf("[B200] geodesic kernels %.3f ms, radiation kernels %.3f ms, %lld rays, %lld samples\n",
                T.gpu_geodesic_ms, T.gpu_radiation_ms, T.rays, T.samples);
  }
  return T;
}

}  // namespace blh
