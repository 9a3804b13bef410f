#include "npz_writer.hpp"

#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>

#include "input_file.hpp"

namespace blh {

namespace {

// slice-by-8 CRC-32 (IEEE 802.3 polynomial, reflected): the payloads reach hundreds of MB
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
      for (int s = 1; s < 8; s++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};

std::vector<uint8_t> npy_with_header(const char *descr, const std::vector<int> &shape, const void *data, size_t bytes) {
  const size_t header_length = 128;
  std::vector<uint8_t> out(header_length + bytes);
  std::memcpy(out.data(), "\x93NUMPY\x01\x00", 8);
  uint16_t hlen = (uint16_t)(header_length - 10);
  std::memcpy(out.data() + 8, &hlen, 2);
  std::string dict = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (";
  for (size_t i = 0; i < shape.size(); i++) {
    dict += std::to_string(shape[i]);
    if (shape.size() == 1) dict += ",";
    else if (i + 1 < shape.size()) dict += ", ";
  }
  dict += ")}";
  if (dict.size() > header_length - 11) throw Error("Error converting data to .npy format.");
  std::memset(out.data() + 10, ' ', header_length - 11);
  std::memcpy(out.data() + 10, dict.data(), dict.size());
  out[header_length - 1] = '\n';
  if (bytes) std::memcpy(out.data() + header_length, data, bytes);
  return out;
}

size_t count_of(const std::vector<int> &shape) {
  size_t n = 1;
  for (int v : shape) n *= (size_t)v;
  return n;
}

template <typename T>
void put(std::vector<uint8_t> &v, T x) {
  uint8_t b[sizeof(T)];
  std::memcpy(b, &x, sizeof(T));
  v.insert(v.end(), b, b + sizeof(T));
}

}  // namespace

namespace {

uint32_t crc32_update(uint32_t c, const uint8_t *p, size_t n) {
  static const CrcTables T;
  while (n >= 8) {
    uint32_t lo, hi;
    std::memcpy(&lo, p, 4);
    std::memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = T.t[7][lo & 0xff] ^ T.t[6][(lo >> 8) & 0xff] ^ T.t[5][(lo >> 16) & 0xff] ^ T.t[4][lo >> 24] ^
        T.t[3][hi & 0xff] ^ T.t[2][(hi >> 8) & 0xff] ^ T.t[1][(hi >> 16) & 0xff] ^ T.t[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = T.t[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return c;
}

// CRC of a concatenation from the CRCs of its parts: crc(A || B) = shift(crc(A), 8 |B| zero bits) ^ crc(B), the shift
// being multiplication by x^(8 |B|) modulo the CRC polynomial, done as a 32 x 32 bit matrix over GF(2) raised to that
// power by repeated squaring (the construction zlib's crc32_combine uses).
uint32_t gf2_times(const uint32_t *mat, uint32_t vec) {
  uint32_t sum = 0;
  for (int i = 0; vec; vec >>= 1, i++)
    if (vec & 1) sum ^= mat[i];
  return sum;
}
void gf2_square(uint32_t *square, const uint32_t *mat) {
  for (int n = 0; n < 32; n++) square[n] = gf2_times(mat, mat[n]);
}
uint32_t crc32_concat(uint32_t crc_a, uint32_t crc_b, size_t len_b) {
  if (len_b == 0) return crc_a;
  uint32_t even[32], odd[32];
  odd[0] = 0xedb88320u;   // one zero bit
  uint32_t row = 1;
  for (int n = 1; n < 32; n++) {
    odd[n] = row;
    row <<= 1;
  }
  gf2_square(even, odd);   // two zero bits
  gf2_square(odd, even);   // four
  do {                      // first pass: one zero byte (eight zero bits), then squares of it
    gf2_square(even, odd);
    if (len_b & 1) crc_a = gf2_times(even, crc_a);
    len_b >>= 1;
    if (len_b == 0) break;
    gf2_square(odd, even);
    if (len_b & 1) crc_a = gf2_times(odd, crc_a);
    len_b >>= 1;
  } while (len_b != 0);
  return crc_a ^ crc_b;
}

}  // namespace

// A 4096^2 multi-frequency frame is gigabytes of payload: the chunks are summed on all host threads and combined.
uint32_t crc32(const uint8_t *p, size_t n) {
  const size_t chunk = (size_t)4 << 20;
  if (n < 2 * chunk) return crc32_update(0xffffffffu, p, n) ^ 0xffffffffu;
  const size_t parts = (n + chunk - 1) / chunk;
  std::vector<uint32_t> crc(parts);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)parts; i++) {
    const size_t at = (size_t)i * chunk, len = at + chunk <= n ? chunk : n - at;
    crc[(size_t)i] = crc32_update(0xffffffffu, p + at, len) ^ 0xffffffffu;
  }
  uint32_t total = crc[0];
  for (size_t i = 1; i < parts; i++) {
    const size_t at = i * chunk, len = at + chunk <= n ? chunk : n - at;
    total = crc32_concat(total, crc[i], len);
  }
  return total;
}

std::vector<uint8_t> npy_bytes(const double *data, const std::vector<int> &shape) {
  return npy_with_header("<f8", shape, data, count_of(shape) * sizeof(double));
}
std::vector<uint8_t> npy_bytes(const int32_t *data, const std::vector<int> &shape) {
  return npy_with_header("<i4", shape, data, count_of(shape) * sizeof(int32_t));
}

void NpzWriter::add(const std::string &name, std::vector<uint8_t> npy) {
  if (npy.size() > 0xffffffffull) throw Error("Array and metadata too large for ZIP record.");
  Entry e;
  e.name = name + ".npy";
  if (e.name.size() > 0xffff) throw Error("Array name too long for ZIP format.");
  e.crc = crc32(npy.data(), npy.size());
  e.data = std::move(npy);
  entries_.push_back(std::move(e));
}

void NpzWriter::write(const std::string &path) const {
  std::ofstream out(path, std::ios::binary);
  if (!out.is_open()) throw Error("Could not open output file.");
  std::time_t now = std::time(nullptr);
  std::tm *lt = std::localtime(&now);
  uint16_t dos_time = (uint16_t)((lt->tm_hour << 11) | ((lt->tm_min & 0x3f) << 5) | ((lt->tm_sec / 2) & 0x1f));
  uint16_t dos_date = (uint16_t)(((lt->tm_year - 80) << 9) | (((lt->tm_mon + 1) & 0xf) << 5) | (lt->tm_mday & 0x1f));
  std::vector<uint8_t> central;
  uint64_t offset = 0;
  for (const Entry &e : entries_) {
    if (offset > 0xffffffffull) throw Error("File too large for ZIP format.");
    std::vector<uint8_t> local;
    put<uint32_t>(local, 0x04034b50u);
    put<uint16_t>(local, 20);           // version needed: 2.0
    put<uint16_t>(local, 0);            // flags
    put<uint16_t>(local, 0);            // stored
    put<uint16_t>(local, dos_time);
    put<uint16_t>(local, dos_date);
    put<uint32_t>(local, e.crc);
    put<uint32_t>(local, (uint32_t)e.data.size());
    put<uint32_t>(local, (uint32_t)e.data.size());
    put<uint16_t>(local, (uint16_t)e.name.size());
    put<uint16_t>(local, 0);
    local.insert(local.end(), e.name.begin(), e.name.end());
    out.write(reinterpret_cast<const char *>(local.data()), (std::streamsize)local.size());
    out.write(reinterpret_cast<const char *>(e.data.data()), (std::streamsize)e.data.size());
    put<uint32_t>(central, 0x02014b50u);
    put<uint16_t>(central, (3 << 8) | 20);  // made by: Unix, 2.0 (what NumPy writes)
    put<uint16_t>(central, 20);
    put<uint16_t>(central, 0);
    put<uint16_t>(central, 0);
    put<uint16_t>(central, dos_time);
    put<uint16_t>(central, dos_date);
    put<uint32_t>(central, e.crc);
    put<uint32_t>(central, (uint32_t)e.data.size());
    put<uint32_t>(central, (uint32_t)e.data.size());
    put<uint16_t>(central, (uint16_t)e.name.size());
    put<uint16_t>(central, 0);   // extra
    put<uint16_t>(central, 0);   // comment
    put<uint16_t>(central, 0);   // disk number
    put<uint16_t>(central, 0);   // internal attributes
    put<uint32_t>(central, 0x81800000u);  // external attributes: regular file, rw-------
    put<uint32_t>(central, (uint32_t)offset);
    central.insert(central.end(), e.name.begin(), e.name.end());
    offset += local.size() + e.data.size();
  }
  if (offset > 0xffffffffull || entries_.size() > 0xffff) throw Error("File too large for ZIP format.");
  out.write(reinterpret_cast<const char *>(central.data()), (std::streamsize)central.size());
  std::vector<uint8_t> end;
  put<uint32_t>(end, 0x06054b50u);
  put<uint16_t>(end, 0);
  put<uint16_t>(end, 0);
  put<uint16_t>(end, (uint16_t)entries_.size());
  put<uint16_t>(end, (uint16_t)entries_.size());
  put<uint32_t>(end, (uint32_t)central.size());
  put<uint32_t>(end, (uint32_t)offset);
  put<uint16_t>(end, 0);
  out.write(reinterpret_cast<const char *>(end.data()), (std::streamsize)end.size());
  if (!out.good()) throw Error("Could not write output file.");
}

}  // namespace blh
