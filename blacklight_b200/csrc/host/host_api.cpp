// C entry points of the host layer (declared in include/blacklight_b200_host.h): what the Python
// tests, bench.py and the command-line driver use to reach the input surface, the camera and the
// re-hosted main without linking C++ types.
#include "../../../include/blacklight_b200_host.h"

#include <cstring>
#include <memory>
#include <string>

#include "config.hpp"
#include "driver.hpp"
#include "npz_writer.hpp"
#include "snapshot.hpp"

struct blh_config {
  blh::RunConfig cfg;
};

struct blh_snapshot {
  blh::RunConfig cfg;                            // the reader refers to it
  std::unique_ptr<blh::SnapshotReader> reader;   // keeps the layout found in the first file (time series)
  blh::AthenaGrid grid;
  double plasma_gamma = 0.0;
};

namespace {
thread_local std::string g_error;
int fail(const std::exception &e) {
  g_error = e.what();
  return 1;
}
}  // namespace

extern "C" {

const char *blh_last_error(void) { return g_error.c_str(); }

int blh_config_from_input(const char *path, blh_config **out) {
  if (!path || !out) { g_error = "null argument"; return 1; }
  *out = nullptr;
  try {
    blh::InputFile in(path);
    blh_config *c = new blh_config{blh::make_config(in)};
    *out = c;
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

void blh_config_free(blh_config *c) { delete c; }

const bl_params *blh_config_params(const blh_config *c) { return c ? &c->cfg.params : nullptr; }

int blh_config_num_runs(const blh_config *c) { return c ? c->cfg.num_runs : 0; }

void blh_config_set_device(blh_config *c, int device, int64_t tile_rays) {
  if (!c) return;
  c->cfg.params.device = device;
  c->cfg.params.tile_rays = tile_rays;
}

void blh_config_set_level0_block_major(blh_config *c, int block_major) {
  if (c) c->cfg.params.level0_block_major = block_major != 0;
}

int blh_camera_frame(const blh_config *c, double out[28]) {
  if (!c || !out) { g_error = "null argument"; return 1; }
  const blh::CameraFrame &f = c->cfg.frame;
  const double *v[7] = {f.x, f.u_con, f.u_cov, f.norm_con, f.norm_con_c, f.hor_con_c, f.vert_con_c};
  for (int i = 0; i < 7; i++) std::memcpy(out + 4 * i, v[i], 4 * sizeof(double));
  return 0;
}

int blh_camera_struct(const blh_config *c, bl_camera *out) {
  if (!c || !out) { g_error = "null argument"; return 1; }
  *out = blh::make_bl_camera(c->cfg);
  return 0;
}

int64_t blh_camera_root(const blh_config *c, double *pos, double *dir, double *factor) {
  if (!c || !pos || !dir || !factor) { g_error = "null argument"; return -1; }
  std::vector<double> p, d, f;
  blh::camera_root(c->cfg.camera, c->cfg.frame, p, d, f);
  std::memcpy(pos, p.data(), p.size() * sizeof(double));
  std::memcpy(dir, d.data(), d.size() * sizeof(double));
  std::memcpy(factor, f.data(), f.size() * sizeof(double));
  return (int64_t)f.size();
}

int64_t blh_camera_refined(const blh_config *c, int level, const int32_t *parent_locs, const uint8_t *flags,
                           int64_t num_parents, int32_t *child_locs, double *pos, double *dir, double *factor) {
  if (!c || !parent_locs || !flags) { g_error = "null argument"; return -1; }
  std::vector<int32_t> pl(parent_locs, parent_locs + 2 * num_parents), cl;
  std::vector<uint8_t> fl(flags, flags + num_parents);
  std::vector<double> p, d, f;
  blh::camera_refined(c->cfg.camera, c->cfg.frame, level, c->cfg.params.adaptive_block_size, pl, fl, cl, p, d, f);
  if (child_locs) std::memcpy(child_locs, cl.data(), cl.size() * sizeof(int32_t));
  if (pos) std::memcpy(pos, p.data(), p.size() * sizeof(double));
  if (dir) std::memcpy(dir, d.data(), d.size() * sizeof(double));
  if (factor) std::memcpy(factor, f.data(), f.size() * sizeof(double));
  return (int64_t)cl.size() / 2;
}

int64_t blh_camera_blocks(const blh_config *c, int level, const int32_t *locs, int64_t num_blocks, double *pos, double *dir,
                          double *factor) {
  if (!c || !locs || !pos || !dir || !factor || num_blocks < 0 || level < 0) { g_error = "bad argument"; return -1; }
  blh::camera_blocks(c->cfg.camera, c->cfg.frame, level, c->cfg.params.adaptive_block_size, locs, num_blocks, pos, dir, factor);
  return num_blocks;
}

int64_t blh_camera_rows(const blh_config *c, const int64_t *rows, int64_t num_rows, double *pos, double *dir, double *factor) {
  if (!c || !rows || !pos || !dir || !factor || num_rows < 0) { g_error = "bad argument"; return -1; }
  for (int64_t r = 0; r < num_rows; r++)
    if (rows[r] < 0 || rows[r] >= c->cfg.camera.resolution) { g_error = "row outside the image"; return -1; }
  std::vector<long long> list(rows, rows + num_rows);
  blh::camera_rows(c->cfg.camera, c->cfg.frame, list.data(), (long long)num_rows, pos, dir, factor);
  return num_rows * c->cfg.camera.resolution;
}

int blh_snapshot_read(const blh_config *c, const char *file, blh_snapshot **out) {
  if (!c || !out) { g_error = "null argument"; return 1; }
  *out = nullptr;
  try {
    if (c->cfg.params.model_type != BL_MODEL_SIMULATION) throw blh::Error("model_type is not simulation.");
    std::unique_ptr<blh_snapshot> s(new blh_snapshot);
    s->cfg = c->cfg;
    s->reader.reset(new blh::SnapshotReader(s->cfg));
    s->reader->read(file ? std::string(file) : s->reader->first_file(), false, s->grid);
    s->plasma_gamma = s->reader->plasma_gamma();
    *out = s.release();
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int blh_snapshot_reread(blh_snapshot *s, const char *file) {
  if (!s || !file) { g_error = "null argument"; return 1; }
  try {
    s->reader->read(file, true, s->grid);
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int blh_snapshot_view(const blh_snapshot *s, bl_grid_view *view, double *time, double *plasma_gamma) {
  if (!s || !view) { g_error = "null argument"; return 1; }
  *view = s->grid.view();
  if (time) *time = s->grid.time;
  if (plasma_gamma) *plasma_gamma = s->plasma_gamma;
  return 0;
}

void blh_snapshot_free(blh_snapshot *s) { delete s; }

int blh_run_input_file(const char *path, int device, int quiet, double timings[12]) {
  if (!path) { g_error = "null argument"; return 1; }
  try {
    blh::RunTimings t = blh::run_input_file(path, device, quiet != 0);
    if (timings) {
      double v[12] = {t.total, t.geodesic, t.read, t.sample, t.image, t.render, t.gpu_geodesic_ms,
                      t.gpu_radiation_ms, t.gpu_refine_ms, (double)t.rays, (double)t.samples, 0.0};
      std::memcpy(timings, v, sizeof v);
    }
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

uint32_t blh_crc32(const void *data, uint64_t bytes) { return blh::crc32(static_cast<const uint8_t *>(data), (size_t)bytes); }

int blh_run_input_file_devices(const char *path, const int *devices, int num_devices, int quiet, double timings[12]) {
  if (!path || !devices || num_devices <= 0) { g_error = "bad argument"; return 1; }
  try {
    blh::RunTimings t = blh::run_input_file(path, std::vector<int>(devices, devices + num_devices), quiet != 0);
    if (timings) {
      double v[12] = {t.total, t.geodesic, t.read, t.sample, t.image, t.render, t.gpu_geodesic_ms,
                      t.gpu_radiation_ms, t.gpu_refine_ms, (double)t.rays, (double)t.samples, (double)num_devices};
      std::memcpy(timings, v, sizeof v);
    }
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

}  // extern "C"
