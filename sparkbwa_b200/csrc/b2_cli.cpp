// `b200bwa` command-line twin of `bwa` (mem only), so the oracle and the engine are driven by
// identical command lines.  "-f <out>" is honoured the way SparkBWA's glue handles it.
#include <string.h>
#include <vector>
#include "../../include/b200bwa.h"
int main(int argc, char** argv) {
  std::vector<char*> av;
  const char* out = 0;
  for (int i = 0; i < argc; ++i) {
    if (i >= 2 && strcmp(argv[i], "-f") == 0 && i + 1 < argc) { out = argv[++i]; continue; }
    av.push_back(argv[i]);
  }
  return b200bwa_main_to((int)av.size(), av.data(), out);
}
