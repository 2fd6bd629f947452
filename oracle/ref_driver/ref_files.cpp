// oracle/ref_driver/ref_files.cpp -- TEST INFRASTRUCTURE ONLY.
// Prints what the reference's own get_file_sample / get_sample_window (include/utilities/files.h:39-110) return for
// the directories given on the command line, one "symbol<TAB>md path<TAB>tas path" line per tuple.  The header is
// included where it lies under /root/reference; it relies on its includer for <tuple>, <stdexcept>, <unistd.h> and
// `using namespace std` (main.cpp provides them upstream).
//   ref_files sample <md_dir> <tas_dir> <symbol>...
//   ref_files window <md_dir> <tas_dir> <symbol> <pattern>...
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unistd.h>
#include <vector>
using namespace std;
#include "utilities/files.h"

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: ref_files sample|window md_dir tas_dir symbol [more...]\n"); return 2; }
  try {
    vector<tuple<string, string, string>> out;
    if (!strcmp(argv[1], "sample")) {
      vector<string> symbols(argv + 4, argv + argc);
      out = get_file_sample(argv[2], argv[3], symbols);
    } else {
      vector<string> patterns(argv + 5, argv + argc);
      out = get_sample_window(argv[2], argv[3], argv[4], patterns);
    }
    for (auto& t : out) printf("%s\t%s\t%s\n", get<0>(t).c_str(), get<1>(t).c_str(), get<2>(t).c_str());
  } catch (const std::exception& e) {
    printf("ERROR\t%s\n", e.what());
    return 1;
  }
  return 0;
}
