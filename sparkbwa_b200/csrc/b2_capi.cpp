// C-ABI entry points (include/b200bwa.h).  b200bwa_main replaces bwa's main() for the `mem`
// sub-command (bwa/main.c:61-104): it builds the @PG line from argv (main.c:68-70) and dispatches
// to the main_mem twin; every other sub-command is reported as unrecognised with return code 1.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/b200bwa.h"
#include "b2_host.h"

static std::string make_pg(int argc, char** argv) {
  std::string pg = "@PG\tID:bwa\tPN:bwa\tVN:" B200BWA_VERSION "\tCL:";
  pg += argc > 0 ? argv[0] : "bwa";
  for (int i = 1; i < argc; ++i) { pg += ' '; pg += argv[i]; }
  return pg;
}

extern "C" {

const char* b200bwa_last_error(void) { return b2_tls_error.c_str(); }

int b200bwa_mem(int argc, char** argv, const char* out_path) {
  try {
    b2_tls_error.clear();
    std::vector<char*> full;
    std::string prog = "bwa";
    full.push_back(&prog[0]);
    for (int i = 0; i < argc; ++i) full.push_back(argv[i]);
    std::string pg = make_pg((int)full.size(), full.data());
    return b2_mem_main(argc, argv, out_path, pg.c_str());
  } catch (const std::exception& e) {
    b2_tls_error = e.what();
    fprintf(stderr, "[E::b200bwa] %s\n", e.what());
    return 1;
  }
}

int b200bwa_main_to(int argc, char** argv, const char* out_path) {
  try {
    b2_tls_error.clear();
    if (argc < 2) {
      fprintf(stderr, "\nProgram: b200bwa (B200-native BWA-MEM behind the SparkBWA bridge)\nVersion: %s\n\nUsage:   bwa mem [options] <idxbase> <in1.fq> [in2.fq]\n\n", B200BWA_VERSION);
      return 1;
    }
    if (strcmp(argv[1], "mem") != 0) {
      fprintf(stderr, "[main] unrecognized command '%s'\n", argv[1]);
      b2_tls_error = std::string("unrecognized command '") + argv[1] + "' (only `mem` is provided by b200bwa)";
      return 1;
    }
    std::string pg = make_pg(argc, argv);
    return b2_mem_main(argc - 1, argv + 1, out_path, pg.c_str());
  } catch (const std::exception& e) {
    b2_tls_error = e.what();
    fprintf(stderr, "[E::b200bwa] %s\n", e.what());
    return 1;
  }
}

int b200bwa_main(int argc, char** argv) { return b200bwa_main_to(argc, argv, 0); }

}  // extern "C"
