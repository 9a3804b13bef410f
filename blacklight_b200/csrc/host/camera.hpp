// Camera frame and per-pixel initial conditions (host FP64, same operation order as the reference so
// the arrays handed to the GPU are bit-identical to reference src/geodesic_integrator/camera.cpp).
#pragma once
#include <cstdint>
#include <vector>

namespace blh {

struct CameraSetup {
  // inputs (radians; geodesic_integrator.cpp:37-52)
  int type = 0;            // 0 plane, 1 pinhole
  double r = 0, th = 0, ph = 0, urn = 0, uthn = 0, uphn = 0, k_r = 0, k_th = 0, k_ph = 0;
  double rotation = 0, width = 0;
  int resolution = 0;
  bool pole = false;
  bool flat = false;
  double a = 0;            // black-hole spin (mass is 1)
  int normalization = 0;   // 0 camera, 1 infinity
};

struct CameraFrame {
  double x[4];                       // camera position (cam_x)
  double u_con[4], u_cov[4];         // camera 4-velocity
  double norm_con[4], norm_con_c[4]; // outward normal (coordinate / camera frame)
  double hor_con_c[4], vert_con_c[4];
};

// camera.cpp:53-380
CameraFrame build_camera_frame(const CameraSetup &s);

// One pixel at fractional image-plane coordinates (u_ind, v_ind) in [-1/2, 1/2] (camera.cpp:528-671).
// pos[4], dir[4] (covariant momentum), *factor = 1 / nu_local.
void camera_pixel(const CameraSetup &s, const CameraFrame &f, double u_ind, double v_ind, double pos[4],
                  double dir[4], double *factor);

// Level-0 camera: pixel m = row * res + col, rows bottom to top (camera.cpp:388-413).
void camera_root(const CameraSetup &s, const CameraFrame &f, std::vector<double> &pos, std::vector<double> &dir,
                 std::vector<double> &factor);

// Refined level: for each flagged parent block, in parent order, its four children
// (2v..2v+1) x (2u..2u+1) at effective resolution res * 2^level (camera.cpp:426-504).
// parent_locs: (B_old,2) (v,u); flags: (B_old); child_locs out: (B_new,2).
void camera_refined(const CameraSetup &s, const CameraFrame &f, int level, int block_size,
                    const std::vector<int32_t> &parent_locs, const std::vector<uint8_t> &flags,
                    std::vector<int32_t> &child_locs, std::vector<double> &pos, std::vector<double> &dir,
                    std::vector<double> &factor);

// Pixels of the given blocks of a level (locs: (blocks,2) (v,u) at effective resolution res * 2^level), block-major:
// pos, dir (blocks * bs^2, 4), factor (blocks * bs^2).  The same per-pixel expressions as camera_refined; what a rank
// that owns only some blocks of a level calls.
void camera_blocks(const CameraSetup &s, const CameraFrame &f, int level, int block_size, const int32_t *locs, long long blocks,
                   double *pos, double *dir, double *factor);

// Level-0 pixels of the listed image rows only (pixel order within a row as camera_root): what a device that owns a share
// of the frame's rows builds.  pos, dir: (num_rows * res, 4); factor: (num_rows * res).
void camera_rows(const CameraSetup &s, const CameraFrame &f, const long long *rows, long long num_rows, double *pos, double *dir,
                 double *factor);

// Child blocks of the flagged parents: parents in index order x 4 children (2v..2v+1) x (2u..2u+1) (camera.cpp:445-459).
void child_blocks(const std::vector<int32_t> &parent_locs, const std::vector<uint8_t> &flags, std::vector<int32_t> &child_locs);

// Image frequencies (camera.cpp:30-50). spacing: 0 lin_freq, 1 lin_wave, 2 log
std::vector<double> image_frequencies(int num, double single, double start, double end, int spacing);

}  // namespace blh
