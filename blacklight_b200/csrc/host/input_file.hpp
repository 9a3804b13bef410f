// Parameter-file surface of the reference, kept verbatim: `key = value  # comment` lines, all
// whitespace removed, unknown key = error, absence of a key detected when it is first needed
// (reference src/input_reader/input_reader.cpp:72-428, enum_readers.cpp, render_reader.cpp,
// adaptive_reader.cpp).  Table-driven instead of one optional<> member per key.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace blh {

// Error type carrying the reference's message texts (utils/exceptions.hpp:14-29)
struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
void warning(const std::string &message);  // "Warning: ...\n" on stderr, as the reference prints

class InputFile {
 public:
  explicit InputFile(const std::string &path);  // parses; throws Error with the reference's messages

  bool has(const std::string &key) const { return values_.count(key) != 0; }
  // Typed getters; a missing key throws (the reference's std::bad_optional_access surfaces as an
  // error in main, blacklight.cpp:101-105)
  const std::string &str(const std::string &key) const;
  bool flag(const std::string &key) const;
  int integer(const std::string &key) const;
  double real(const std::string &key) const;
  float real32(const std::string &key) const;
  void triple(const std::string &key, double out[3]) const;
  // enumerations: index of the value within `names`, with the reference's error text on mismatch
  int choice(const std::string &key, const std::vector<std::string> &names, const char *type_name) const;

  // values derived at parse time
  bool camera_pole() const { return camera_pole_; }   // camera_th exactly 0 or 180 (input_reader.cpp:492-500)
  int num_runs() const;                               // input_reader.cpp:419-427

 private:
  std::map<std::string, std::string> values_;
  bool camera_pole_ = false;
};

}  // namespace blh
