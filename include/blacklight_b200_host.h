/* blacklight_b200 host layer -- C entry points above the kernel ABI (blacklight_b200.h):
 * the reference's .input parameter surface, its camera construction and its main loop, re-hosted.
 *   blh_config_from_input   <- InputReader::Read + the two integrator constructors
 *                              (input_reader.cpp:72, geodesic_integrator.cpp:23, radiation_integrator.cpp:26)
 *   blh_camera_root/refined <- GeodesicIntegrator::InitializeCamera / AugmentCamera (camera.cpp:27,426)
 *   blh_run_input_file      <- main (blacklight.cpp:31-273)
 * All functions return 0 on success (or a count where stated); blh_last_error() gives the message,
 * which for user errors is the reference's own text. */
#ifndef BLACKLIGHT_B200_HOST_H_
#define BLACKLIGHT_B200_HOST_H_

#include "blacklight_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct blh_config blh_config;

const char *blh_last_error(void);
int blh_config_from_input(const char *path, blh_config **out);
void blh_config_free(blh_config *cfg);
const bl_params *blh_config_params(const blh_config *cfg);   /* owned by cfg; camera frame filled in */
int blh_config_num_runs(const blh_config *cfg);
/* B200 knobs of bl_params: CUDA device ordinal and rays per wave (0 = sized from free HBM) */
void blh_config_set_device(blh_config *cfg, int device, int64_t tile_rays);
/* bl_params.level0_block_major (sharded adaptive runs) */
void blh_config_set_level0_block_major(blh_config *cfg, int block_major);
/* cam_x, u_con, u_cov, norm_con, norm_con_c, hor_con_c, vert_con_c: 7 x 4 doubles */
int blh_camera_frame(const blh_config *cfg, double out[28]);
/* The same frame with the per-pixel parameters, as bl_set_camera takes it (pixels are then generated on the device by
 * bl_trace_level_pixels; the blh_camera_* functions below are the host versions, kept for output_camera, checkpoints
 * and as the parity reference of the device kernel). */
int blh_camera_struct(const blh_config *cfg, bl_camera *out);
/* pos, dir: (res*res,4); factor: (res*res).  Returns the number of rays. */
int64_t blh_camera_root(const blh_config *cfg, double *pos, double *dir, double *factor);
/* Children of the flagged parents.  Output buffers sized for 4 * (#flags set) blocks of
 * adaptive_block_size^2 pixels (any may be NULL to query the count).  Returns the number of child blocks. */
int64_t blh_camera_refined(const blh_config *cfg, int level, const int32_t *parent_locs, const uint8_t *flags,
                           int64_t num_parents, int32_t *child_locs, double *pos, double *dir, double *factor);
/* Pixels of the listed blocks of `level` only (locs: (num_blocks,2) block (v,u) at that level), block-major -- for a rank
 * that owns a share of a level's blocks (AugmentCamera's per-pixel expressions, camera.cpp:461-503). */
int64_t blh_camera_blocks(const blh_config *cfg, int level, const int32_t *locs, int64_t num_blocks, double *pos, double *dir,
                          double *factor);
/* Level-0 pixels of the listed image rows only (what a GPU that owns a share of the frame's rows traces): pos, dir
 * (num_rows * res, 4), factor (num_rows * res), bit-identical to the same rows of blh_camera_root.  Returns the ray count. */
int64_t blh_camera_rows(const blh_config *cfg, const int64_t *rows, int64_t num_rows, double *pos, double *dir, double *factor);
/* timings: total, geodesic, read, sample, image, render [s]; gpu geodesic, radiation, refine [ms];
 * rays, samples, reserved */
/* Snapshot readers -- the upload side of SimulationReader::Read (simulation_reader.cpp:200-861): simulation_format
 * athena (.athdf), athenak (binary dump) or harm3d, as the input file names it.  file = NULL reads the input file's
 * simulation_file (first of the series).  The view's arrays are owned by the snapshot and are what bl_upload_grid takes;
 * plasma_gamma is the adiabatic index the reader settled on (the file's where the input file gives none). */
typedef struct blh_snapshot blh_snapshot;
int blh_snapshot_read(const blh_config *cfg, const char *file, blh_snapshot **out);
/* Next file of a time series: keeps the layout (coordinates, variable positions, adiabatic indices) found in the first
 * file and refreshes the cell data and the time, as the reference's reader does after its first call. */
int blh_snapshot_reread(blh_snapshot *snap, const char *file);
int blh_snapshot_view(const blh_snapshot *snap, bl_grid_view *view, double *time, double *plasma_gamma);
void blh_snapshot_free(blh_snapshot *snap);

/* CRC-32 (IEEE 802.3, as zip / zlib) of a host buffer, the npz writer's checksum (the reference's is a byte-at-a-time loop,
 * zip_format.cpp:289-362): slice-by-8 per chunk on all host threads, chunk sums combined in GF(2). */
uint32_t blh_crc32(const void *data, uint64_t bytes);

int blh_run_input_file(const char *path, int device, int quiet, double timings[12]);
/* The same run spread over several GPUs of the node from this one process (the reference parallelises inside main too,
 * blacklight.cpp:77): one context and one host thread per listed CUDA device, the image rows -- refinement blocks for
 * adaptive runs -- dealt round-robin over them, the grid replicated on each; the written output is bitwise the single
 * device's.  blh_run_input_file(path, -1, ...) and the executable take the list from BLACKLIGHT_DEVICES ("all", "0-7",
 * "0,2,3").  timings[11] = number of devices used. */
int blh_run_input_file_devices(const char *path, const int *devices, int num_devices, int quiet, double timings[12]);

#ifdef __cplusplus
}
#endif
#endif
