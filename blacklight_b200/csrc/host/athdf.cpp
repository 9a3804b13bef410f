#include "athdf.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>

#include "input_file.hpp"

namespace blh {

namespace {

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
    list_group();
  }

  bool has_dataset(const std::string &name) const { return children_.count(name) != 0; }

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
    auto it = children_.find(name);
    if (it == children_.end()) throw Error("Could not find dataset " + name + " in HDF5 file.");
    uint64_t addr = ~0ull, bytes = 0;
    bool have_dt = false, have_ds = false, have_layout = false;
    for (const Message &m : messages(it->second)) {
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

  // walk the root group's B-tree (v1, node type 0) down to its symbol-table nodes
  void list_group() {
    const uint8_t *heap = at(root_heap_, 32);
    if (std::memcmp(heap, "HEAP", 4) != 0) throw Error("Unexpected HDF5 heap signature.");
    uint64_t heap_data = rd64(heap + 24);
    std::vector<uint64_t> nodes = {root_btree_};
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
          children_[name] = header;
        }
      } else {
        throw Error("Unexpected HDF5 group node signature.");
      }
    }
  }
};

std::vector<std::string> string_attribute(const H5File &f, const std::string &name) {
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

std::vector<int32_t> int_values(const uint8_t *d, const Datatype &dt, size_t count, const std::string &what) {
  if (dt.cls != 0 || (dt.size != 4 && dt.size != 8)) throw Error("Unexpected HDF5 integer type for " + what + ".");
  std::vector<int32_t> out(count);
  for (size_t i = 0; i < count; i++) std::memcpy(&out[i], d + i * dt.size, 4);  // low 4 bytes of little-endian value
  return out;
}

std::vector<int32_t> int_attribute(const H5File &f, const std::string &name) {
  Datatype dt;
  size_t count = 0;
  const uint8_t *d = f.attribute(name, dt, count);
  return int_values(d, dt, count, name);
}

void coordinate_dataset(const H5File &f, const std::string &name, int n_b, std::vector<double> &out, int &n) {
  Datatype dt;
  std::vector<uint64_t> dims;
  const uint8_t *d = f.dataset(name, dt, dims);
  if (dt.cls != 1 || dt.size != 4 || dims.size() != 2 || (int)dims[0] != n_b) throw Error("Unexpected layout of dataset " + name + ".");
  n = (int)dims[1];
  out.resize((size_t)n_b * n);
  for (size_t i = 0; i < out.size(); i++) {
    float v;
    std::memcpy(&v, d + 4 * i, 4);
    out[i] = static_cast<double>(v);
  }
}

int find_name(const std::vector<std::string> &names, int begin, int end, const std::string &want) {
  for (int i = begin; i < end; i++)
    if (names[(size_t)i] == want) return i;
  return -1;
}

}  // namespace

bl_grid_view AthenaGrid::view() const {
  bl_grid_view v{};
  v.n_b = n_b; v.n_k = n_k; v.n_j = n_j; v.n_i = n_i; v.n_var = n_var;
  v.levels = levels.data(); v.locations = locations.data();
  v.x1f = x1f.data(); v.x2f = x2f.data(); v.x3f = x3f.data();
  v.x1v = x1v.data(); v.x2v = x2v.data(); v.x3v = x3v.data();
  v.prim = prim.data();
  v.ind_rho = ind_rho; v.ind_pgas = ind_pgas; v.ind_kappa = ind_kappa;
  v.ind_uu1 = ind_uu1; v.ind_uu2 = ind_uu2; v.ind_uu3 = ind_uu3;
  v.ind_bb1 = ind_bb1; v.ind_bb2 = ind_bb2; v.ind_bb3 = ind_bb3;
  v.n_3_root = n_3_root;
  return v;
}

double read_athdf_time(const std::string &path) {
  H5File f(path);
  Datatype dt;
  size_t count = 0;
  const uint8_t *d = f.attribute("Time", dt, count);
  if (dt.cls != 1 || dt.size != 4) throw Error("Unexpected HDF5 datatype for attribute Time.");
  float t;
  std::memcpy(&t, d, 4);
  return t;
}

void read_athdf(const std::string &path, const std::string &kappa_name, bool reuse_layout, AthenaGrid &g) {
  H5File f(path);
  {
    Datatype dt;
    size_t count = 0;
    const uint8_t *d = f.attribute("Time", dt, count);
    if (dt.cls != 1 || dt.size != 4) throw Error("Unexpected HDF5 datatype for attribute Time.");
    float t;
    std::memcpy(&t, d, 4);
    g.time = t;
  }
  if (!reuse_layout) {
    std::vector<int32_t> root = int_attribute(f, "RootGridSize");
    if (root.size() != 3) throw Error("Unexpected RootGridSize in data file.");
    g.n_3_root = root[2];
    Datatype dt;
    std::vector<uint64_t> dims;
    const uint8_t *d = f.dataset("Levels", dt, dims);
    if (dims.size() != 1) throw Error("Unexpected layout of dataset Levels.");
    g.n_b = (int)dims[0];
    g.levels = int_values(d, dt, (size_t)g.n_b, "Levels");
    d = f.dataset("LogicalLocations", dt, dims);
    if (dims.size() != 2 || (int)dims[0] != g.n_b || dims[1] != 3) throw Error("Unexpected layout of dataset LogicalLocations.");
    g.locations = int_values(d, dt, (size_t)g.n_b * 3, "LogicalLocations");
    int n1f, n2f, n3f;
    coordinate_dataset(f, "x1f", g.n_b, g.x1f, n1f);
    coordinate_dataset(f, "x2f", g.n_b, g.x2f, n2f);
    coordinate_dataset(f, "x3f", g.n_b, g.x3f, n3f);
    coordinate_dataset(f, "x1v", g.n_b, g.x1v, g.n_i);
    coordinate_dataset(f, "x2v", g.n_b, g.x2v, g.n_j);
    coordinate_dataset(f, "x3v", g.n_b, g.x3v, g.n_k);
    if (n1f != g.n_i + 1 || n2f != g.n_j + 1 || n3f != g.n_k + 1) throw Error("Inconsistent face and cell coordinate arrays.");

    // variable bookkeeping: datasets are stacked "prim" then "B"; indices refer to the stacked array
    std::vector<std::string> dataset_names = string_attribute(f, "DatasetNames");
    std::vector<std::string> variable_names = string_attribute(f, "VariableNames");
    std::vector<int32_t> num_variables = int_attribute(f, "NumVariables");
    if (num_variables.size() != dataset_names.size()) throw Error("Inconsistent dataset metadata in data file.");
    int ind_hydro = -1, ind_bb = -1, prim_off = 0, bb_off = 0, running = 0;
    for (size_t i = 0; i < dataset_names.size(); i++) {
      if (dataset_names[i] == "prim" && ind_hydro < 0) { ind_hydro = (int)i; prim_off = running; }
      if (dataset_names[i] == "B" && ind_bb < 0) { ind_bb = (int)i; bb_off = running; }
      running += num_variables[i];
    }
    if (ind_hydro < 0) throw Error("Unable to locate array \"prim\" in data file.");
    if (ind_bb < 0) throw Error("Unable to locate array \"B\" in data file.");
    int n_hydro = num_variables[(size_t)ind_hydro], n_bb = num_variables[(size_t)ind_bb];
    auto hydro = [&](const char *nm, const char *msg) {
      int i = find_name(variable_names, prim_off, prim_off + n_hydro, nm);
      if (i < 0) throw Error(msg);
      return i - prim_off;
    };
    g.ind_rho = hydro("rho", "Unable to locate \"rho\" slice of \"prim\" in data file.");
    g.ind_pgas = hydro("press", "Unable to locate \"press\" slice of \"prim\" in data file.");
    if (!kappa_name.empty()) g.ind_kappa = hydro(kappa_name.c_str(), "Unable to locate electron entropy slice of \"prim\" in data file.");
    g.ind_uu1 = hydro("vel1", "Unable to locate \"vel1\" slice of \"prim\" in data file.");
    g.ind_uu2 = hydro("vel2", "Unable to locate \"vel2\" slice of \"prim\" in data file.");
    g.ind_uu3 = hydro("vel3", "Unable to locate \"vel3\" slice of \"prim\" in data file.");
    auto field = [&](const char *nm, const char *msg) {
      int i = find_name(variable_names, bb_off, bb_off + n_bb, nm);
      if (i < 0) throw Error(msg);
      return n_hydro + (i - bb_off);
    };
    g.ind_bb1 = field("Bcc1", "Unable to locate \"Bcc1\" slice of \"prim\" in data file.");
    g.ind_bb2 = field("Bcc2", "Unable to locate \"Bcc2\" slice of \"prim\" in data file.");
    g.ind_bb3 = field("Bcc3", "Unable to locate \"Bcc3\" slice of \"prim\" in data file.");
    g.n_var = n_hydro + n_bb;
    g.prim.resize((size_t)g.n_var * g.n_b * g.n_k * g.n_j * g.n_i);
  }
  size_t cells = (size_t)g.n_b * g.n_k * g.n_j * g.n_i;
  size_t filled = 0;
  for (const char *name : {"prim", "B"}) {
    Datatype dt;
    std::vector<uint64_t> dims;
    const uint8_t *d = f.dataset(name, dt, dims);
    if (dt.cls != 1 || dt.size != 4 || dims.size() != 5 || (int)dims[1] != g.n_b || (int)dims[2] != g.n_k ||
        (int)dims[3] != g.n_j || (int)dims[4] != g.n_i)
      throw Error(std::string("Unexpected layout of dataset ") + name + ".");
    size_t n = (size_t)dims[0] * cells;
    if (filled + n > g.prim.size()) throw Error("Cell data larger than declared number of variables.");
    std::memcpy(g.prim.data() + filled, d, n * sizeof(float));
    filled += n;
  }
  if (filled != g.prim.size()) throw Error("Cell data smaller than declared number of variables.");
}

std::string format_numbered(const std::string &pattern, int number, const char *what) {
  std::string err = std::string("Invalid ") + what + " for multiple runs.";
  std::string::size_type open = pattern.find_first_of('{');
  if (open == std::string::npos) throw Error(err);
  std::string::size_type close = pattern.find_first_of('}', open);
  if (close == std::string::npos) throw Error(err);
  if (pattern[close - 1] != 'd') throw Error(err);
  int width = 0;
  if (close - open > 2) width = std::stoi(pattern.substr(open + 1, close - open - 2));
  char digits[32];
  int len = std::snprintf(digits, sizeof digits, "%d", number);
  std::string out = pattern.substr(0, open);
  for (int i = len; i < width; i++) out += '0';
  out += digits;
  out += pattern.substr(close + 1);
  return out;
}

}  // namespace blh
