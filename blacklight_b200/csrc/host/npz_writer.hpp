// Output formats of the reference, byte-compatible in structure: .npy v1.0 with a fixed 128-byte
// header, .npz as a stored (method 0) ZIP 2.0 archive without ZIP64, and raw little-endian dumps
// (reference src/output_writer/numpy_format.cpp:46-744, zip_format.cpp:26-362, raw_format.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace blh {

// One array as .npy bytes.  shape is outermost-first (C order).
std::vector<uint8_t> npy_bytes(const double *data, const std::vector<int> &shape);
std::vector<uint8_t> npy_bytes(const int32_t *data, const std::vector<int> &shape);

class NpzWriter {
 public:
  void add(const std::string &name, std::vector<uint8_t> npy);  // stored as "<name>.npy"
  void write(const std::string &path) const;                    // throws blh::Error

 private:
  struct Entry {
    std::string name;
    std::vector<uint8_t> data;
    uint32_t crc;
  };
  std::vector<Entry> entries_;
};

uint32_t crc32(const uint8_t *data, size_t n);

}  // namespace blh
