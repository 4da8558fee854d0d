// `b200bwa` command-line twin of `bwa` (mem only), so the oracle and the engine are driven by
// identical command lines.  "-f <out>" is honoured the way SparkBWA's glue handles it.
// Test hook: B200BWA_CLI_CONCURRENT=n runs the same command from n threads at once (outputs <out>.0 .. <out>.n-1), the
// way several task threads of one Spark executor call the JNI symbol concurrently.
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>
#include "../../include/b200bwa.h"
int main(int argc, char** argv) {
  std::vector<char*> av;
  const char* out = 0;
  for (int i = 0; i < argc; ++i) {
    if (i >= 2 && strcmp(argv[i], "-f") == 0 && i + 1 < argc) { out = argv[++i]; continue; }
    av.push_back(argv[i]);
  }
  const char* cc = getenv("B200BWA_CLI_CONCURRENT");
  const int n = cc ? atoi(cc) : 0;
  if (n > 1 && out) {
    std::vector<int> rc(n, 0);
    std::vector<std::thread> th;
    for (int k = 0; k < n; ++k)
      th.emplace_back([&, k]() { const std::string o = std::string(out) + "." + std::to_string(k); rc[k] = b200bwa_main_to((int)av.size(), av.data(), o.c_str()); });
    for (auto& t : th) t.join();
    for (int k = 0; k < n; ++k) if (rc[k]) return rc[k];
    return 0;
  }
  return b200bwa_main_to((int)av.size(), av.data(), out);
}
