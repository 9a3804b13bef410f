// blacklight_b200 <file.input> -- drop-in for the reference's `bin/blacklight <file.input>`
// (reference src/blacklight.cpp:31-47): same parameter file, same outputs, same timing report.
#include <cstdio>

#include "../../../include/blacklight_b200_host.h"

int main(int argc, char *argv[]) {
  if (argc != 2) {
    std::printf("Error: Must give a single input file.\n");
    return 1;
  }
  if (blh_run_input_file(argv[1], -1, 0, nullptr) != 0) {
    std::printf("Error: %s\n", blh_last_error());
    return 1;
  }
  return 0;
}
