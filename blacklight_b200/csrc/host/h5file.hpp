// The subset of HDF5 the reference's hand-written parser accepts (reference src/simulation_reader/hdf5_format_*.cpp):
// superblock v0/v1, old-style groups (symbol table, local heap, v1 B-tree), v1 object headers, contiguous
// datasets of little-endian integers / IEEE floats / fixed-length strings, v1 attributes on the root group.
// Shared by the Athena++ (.athdf) and iharm3d readers.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "input_file.hpp"

namespace blh {
namespace h5 {

struct Message {
  int type;
  const uint8_t *data;
  size_t size;
};

struct Datatype {
  int cls = -1;      // 0 fixed point, 1 float, 3 string
  uint32_t size = 0; // bytes per element
};

class H5File {
 public:
  explicit H5File(const std::string &path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f.is_open()) throw Error("Could not open file for reading.");
    std::streamsize n = f.tellg();
    f.seekg(0);
    buf_.resize((size_t)n);
    if (!f.read(reinterpret_cast<char *>(buf_.data()), n)) throw Error("Could not read simulation file.");
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (buf_.size() < 96 || std::memcmp(buf_.data(), sig, 8) != 0) throw Error("Unexpected HDF5 format signature.");
    int version = buf_[8];
    if (version != 0 && version != 1) throw Error("Unexpected HDF5 superblock version.");
    if (buf_[13] != 8 || buf_[14] != 8) throw Error("Unexpected HDF5 size of offsets or lengths.");
    size_t pos = version == 0 ? 24 : 28;   // after K values and flags (v1 adds indexed-storage K + reserved)
    pos += 32;                              // base, free-space, end-of-file, driver-info addresses
    // root group symbol table entry
    root_header_ = u64(pos + 8);
    uint32_t cache = u32(pos + 16);
    if (cache == 1) {
      root_btree_ = u64(pos + 24);
      root_heap_ = u64(pos + 32);
    } else {
      for (const Message &m : messages(root_header_))
        if (m.type == 0x11) {
          root_btree_ = rd64(m.data);
          root_heap_ = rd64(m.data + 8);
        }
    }
    if (!root_btree_ || !root_heap_) throw Error("Unexpected HDF5 root group layout.");
    children_ = list_group(root_btree_, root_heap_);
  }

  bool has_dataset(const std::string &name) const {
    try {
      return resolve(name) != 0;
    } catch (const Error &) {
      return false;
    }
  }

  // attribute on the root group: raw bytes + datatype + element count
  const uint8_t *attribute(const std::string &name, Datatype &dt, size_t &count) const {
    for (const Message &m : messages(root_header_)) {
      if (m.type != 0x0c) continue;
      const uint8_t *d = m.data;
      int version = d[0];
      if (version != 1) throw Error("Unexpected HDF5 attribute message version.");
      size_t name_size = rd16(d + 2), dt_size = rd16(d + 4), ds_size = rd16(d + 6);
      const uint8_t *pn = d + 8;
      std::string attr_name(reinterpret_cast<const char *>(pn));
      const uint8_t *pdt = pn + pad8(name_size);
      const uint8_t *pds = pdt + pad8(dt_size);
      const uint8_t *pdata = pds + pad8(ds_size);
      if (attr_name != name) continue;
      dt = parse_datatype(pdt);
      std::vector<uint64_t> dims = parse_dataspace(pds);
      count = 1;
      for (uint64_t v : dims) count *= (size_t)v;
      return pdata;
    }
    throw Error("Could not find attribute " + name + " in HDF5 file.");
  }

  // dataset: raw bytes + datatype + dims
  const uint8_t *dataset(const std::string &name, Datatype &dt, std::vector<uint64_t> &dims) const {
    uint64_t addr = ~0ull, bytes = 0;
    bool have_dt = false, have_ds = false, have_layout = false;
    for (const Message &m : messages(resolve(name))) {
      if (m.type == 0x03) { dt = parse_datatype(m.data); have_dt = true; }
      else if (m.type == 0x01) { dims = parse_dataspace(m.data); have_ds = true; }
      else if (m.type == 0x08) {
        int version = m.data[0];
        if (version == 3) {
          if (m.data[1] != 1) throw Error("Only contiguous HDF5 datasets are supported.");
          addr = rd64(m.data + 2);
          bytes = rd64(m.data + 10);
        } else if (version == 1 || version == 2) {
          int rank = m.data[1];
          if (m.data[2] != 1) throw Error("Only contiguous HDF5 datasets are supported.");
          addr = rd64(m.data + 8);
          (void)rank;
        } else {
          throw Error("Unexpected HDF5 data layout message version.");
        }
        have_layout = true;
      }
    }
    if (!have_dt || !have_ds || !have_layout || addr == ~0ull) throw Error("Incomplete HDF5 dataset header for " + name + ".");
    size_t need = dt.size;
    for (uint64_t v : dims) need *= (size_t)v;
    if (bytes && bytes < need) throw Error("HDF5 dataset " + name + " is shorter than its dataspace.");
    if (addr + need > buf_.size()) throw Error("HDF5 dataset " + name + " extends past end of file.");
    return buf_.data() + addr;
  }

 private:
  std::vector<uint8_t> buf_;
  uint64_t root_header_ = 0, root_btree_ = 0, root_heap_ = 0;
  std::map<std::string, uint64_t> children_;

  static size_t pad8(size_t n) { return (n + 7) / 8 * 8; }
  static uint16_t rd16(const uint8_t *p) { uint16_t v; std::memcpy(&v, p, 2); return v; }
  static uint32_t rd32(const uint8_t *p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
  static uint64_t rd64(const uint8_t *p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
  const uint8_t *at(size_t pos, size_t n) const {
    if (pos + n > buf_.size()) throw Error("Unexpected end of HDF5 file.");
    return buf_.data() + pos;
  }
  uint32_t u32(size_t pos) const { return rd32(at(pos, 4)); }
  uint64_t u64(size_t pos) const { return rd64(at(pos, 8)); }

  // all messages of a version-1 object header, following continuation blocks
  std::vector<Message> messages(uint64_t addr) const {
    std::vector<Message> out;
    const uint8_t *h = at(addr, 16);
    if (h[0] != 1) throw Error("Unexpected HDF5 object header version.");
    int remaining = rd16(h + 2);
    uint32_t header_size = rd32(h + 8);
    std::vector<std::pair<uint64_t, uint64_t>> blocks = {{addr + 16, header_size}};
    for (size_t b = 0; b < blocks.size() && remaining > 0; b++) {
      uint64_t pos = blocks[b].first, end = blocks[b].first + blocks[b].second;
      while (pos + 8 <= end && remaining > 0) {
        const uint8_t *m = at(pos, 8);
        int type = rd16(m);
        size_t size = rd16(m + 2);
        int flags = m[4];
        const uint8_t *data = at(pos + 8, size);
        remaining--;
        if (type == 0x10) blocks.push_back({rd64(data), rd64(data + 8)});
        else if (!(flags & 0x02)) out.push_back({type, data, size});
        pos += 8 + size;
      }
    }
    return out;
  }

  static Datatype parse_datatype(const uint8_t *p) {
    Datatype dt;
    dt.cls = p[0] & 0x0f;
    dt.size = rd32(p + 4);
    if (dt.cls != 0 && dt.cls != 1 && dt.cls != 3) throw Error("Unexpected HDF5 datatype class.");
    if ((p[1] & 0x01) && dt.cls != 3) throw Error("Big-endian HDF5 data are not supported.");
    return dt;
  }

  static std::vector<uint64_t> parse_dataspace(const uint8_t *p) {
    int version = p[0], rank = p[1];
    const uint8_t *d = version == 1 ? p + 8 : p + 4;
    if (version != 1 && version != 2) throw Error("Unexpected HDF5 dataspace message version.");
    std::vector<uint64_t> dims((size_t)rank);
    for (int r = 0; r < rank; r++) dims[(size_t)r] = rd64(d + 8 * r);
    return dims;
  }

  // object header address of `a/b/c`: groups on the way are old-style (symbol-table message, type 0x11)
  uint64_t resolve(const std::string &path) const {
    const std::map<std::string, uint64_t> *level = &children_;
    std::map<std::string, uint64_t> listed;
    size_t begin = 0;
    for (;;) {
      size_t slash = path.find('/', begin);
      std::string part = path.substr(begin, slash == std::string::npos ? std::string::npos : slash - begin);
      auto it = level->find(part);
      if (it == level->end()) throw Error("Could not find HDF5 dataset in file.");
      if (slash == std::string::npos) return it->second;
      uint64_t btree = 0, heap = 0;
      for (const Message &m : messages(it->second))
        if (m.type == 0x11) {
          btree = rd64(m.data);
          heap = rd64(m.data + 8);
        }
      if (!btree || !heap) throw Error("Could not find HDF5 dataset in file.");
      listed = list_group(btree, heap);
      level = &listed;
      begin = slash + 1;
    }
  }

  // walk a group's B-tree (v1, node type 0) down to its symbol-table nodes
  std::map<std::string, uint64_t> list_group(uint64_t btree, uint64_t heap_addr) const {
    std::map<std::string, uint64_t> children;
    const uint8_t *heap = at(heap_addr, 32);
    if (std::memcmp(heap, "HEAP", 4) != 0) throw Error("Unexpected HDF5 heap signature.");
    uint64_t heap_data = rd64(heap + 24);
    std::vector<uint64_t> nodes = {btree};
    while (!nodes.empty()) {
      uint64_t addr = nodes.back();
      nodes.pop_back();
      const uint8_t *n = at(addr, 24);
      if (std::memcmp(n, "TREE", 4) == 0) {
        if (n[4] != 0) throw Error("Unexpected HDF5 B-tree node type.");
        int entries = rd16(n + 6);
        for (int e = 0; e < entries; e++) nodes.push_back(u64(addr + 24 + 16 * (size_t)e + 8));
      } else if (std::memcmp(n, "SNOD", 4) == 0) {
        int symbols = rd16(n + 6);
        for (int s = 0; s < symbols; s++) {
          size_t entry = addr + 8 + 40 * (size_t)s;
          uint64_t name_off = u64(entry), header = u64(entry + 8);
          const char *name = reinterpret_cast<const char *>(at(heap_data + name_off, 1));
          children[name] = header;
        }
      } else {
        throw Error("Unexpected HDF5 group node signature.");
      }
    }
    return children;
  }
};

inline std::vector<std::string> string_attribute(const H5File &f, const std::string &name) {
  Datatype dt;
  size_t count = 0;
  const uint8_t *d = f.attribute(name, dt, count);
  if (dt.cls != 3) throw Error("Unexpected HDF5 datatype for attribute " + name + ".");
  std::vector<std::string> out;
  for (size_t i = 0; i < count; i++) {
    const char *s = reinterpret_cast<const char *>(d + i * dt.size);
    size_t len = 0;
    while (len < dt.size && s[len] != '\0') len++;
    out.emplace_back(s, len);
  }
  return out;
}

inline std::vector<int32_t> int_values(const uint8_t *d, const Datatype &dt, size_t count, const std::string &what) {
  if (dt.cls != 0 || (dt.size != 4 && dt.size != 8)) throw Error("Unexpected HDF5 integer type for " + what + ".");
  std::vector<int32_t> out(count);
  for (size_t i = 0; i < count; i++) std::memcpy(&out[i], d + i * dt.size, 4);  // low 4 bytes of little-endian value
  return out;
}

inline std::vector<int32_t> int_attribute(const H5File &f, const std::string &name) {
  Datatype dt;
  size_t count = 0;
  const uint8_t *d = f.attribute(name, dt, count);
  return int_values(d, dt, count, name);
}

// scalar or 1-d dataset of 8-byte floats / of integers / of fixed-length strings, by path
inline std::vector<double> double_dataset(const H5File &f, const std::string &path) {
  Datatype dt;
  std::vector<uint64_t> dims;
  const uint8_t *d = f.dataset(path, dt, dims);
  if (dt.cls != 1 || dt.size != 8) throw Error("Unexpected double size.");
  size_t count = 1;
  for (uint64_t v : dims) count *= (size_t)v;
  std::vector<double> out(count);
  std::memcpy(out.data(), d, count * 8);
  return out;
}

inline std::vector<int32_t> int_dataset(const H5File &f, const std::string &path) {
  Datatype dt;
  std::vector<uint64_t> dims;
  const uint8_t *d = f.dataset(path, dt, dims);
  size_t count = 1;
  for (uint64_t v : dims) count *= (size_t)v;
  return int_values(d, dt, count, path);
}

inline std::vector<std::string> string_dataset(const H5File &f, const std::string &path) {
  Datatype dt;
  std::vector<uint64_t> dims;
  const uint8_t *d = f.dataset(path, dt, dims);
  if (dt.cls != 3 || dims.size() > 1) throw Error("Unexpected HDF5 string array size.");
  size_t count = dims.empty() ? 1 : (size_t)dims[0];
  std::vector<std::string> out;
  for (size_t i = 0; i < count; i++) {
    const char *s = reinterpret_cast<const char *>(d + i * dt.size);
    size_t len = 0;
    while (len < dt.size && s[len] != '\0') len++;
    out.emplace_back(s, len);
  }
  return out;
}

}  // namespace h5
}  // namespace blh
