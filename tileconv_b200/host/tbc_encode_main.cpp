// tbc_encode -- minimal command line over the host layer: TIS/MOS -> TBC/MBC on the GPU with
// tileconv's flags for this path (-t 1|2|3, -q <dec><enc> where only the encode digit matters,
// -u, -T, -j n, -s, -o file, -I info only).  See tileconv/options.cpp:48,75-219 for the originals.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tile_stream.h"

int main(int argc, char** argv)
{
  tc::Options options;
  std::string input, output;
  std::vector<std::string> inputs;
  bool silent = false, info = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : ""; };
    if (a == "-t") {
      const int t = std::atoi(next());
      if (t < 0 || t > 3) { std::fprintf(stderr, "Unsupported pixel encoding type\n"); return 1; }
      options.setEncoding(t == 0 ? tc::Encoding::RAW : t == 1 ? tc::Encoding::BC1 : t == 2 ? tc::Encoding::BC2 : tc::Encoding::BC3);
    } else if (a == "-q") {
      const char* q = next();                      // "<decode><encode>", '-' keeps the default (options.cpp:129-152)
      if (std::strlen(q) >= 2 && q[1] >= '0' && q[1] <= '9') options.setEncodingQuality(q[1] - '0');
    } else if (a == "-u") options.setDeflate(false);
    else if (a == "-T") options.setAssumeTis(true);
    else if (a == "-j") options.setThreads(std::atoi(next()));
    else if (a == "-s") silent = true;
    else if (a == "-I") info = true;
    else if (a == "-o") output = next();
    else { input = a; inputs.push_back(a); }
  }
  if (info) {
    // tileconv.cpp:77-81: with -I every input is described and nothing is converted
    for (const std::string& file : inputs) { tc::showInfo(file, options.assumeTis()); std::printf("\n"); }
    return 0;
  }
  if (input.empty()) {
    std::fprintf(stderr, "usage: tbc_encode [-t 0|1|2|3] [-q -N] [-u] [-T] [-j n] [-s] [-I] [-o out] in.tis|in.mos\n");
    return 1;
  }
  const tc::InputKind kind = tc::sniffInput(input, options.assumeTis());
  if (output.empty()) output = input + (kind == tc::InputKind::MOS ? ".mbc" : ".tbc");
  std::string error;
  bool ok = false;
  if (kind == tc::InputKind::TIS) ok = tc::encodeTisFile(options, input, output, error);
  else if (kind == tc::InputKind::MOS) ok = tc::encodeMosFile(options, input, output, error);
  else error = "Unsupported input file";
  if (!ok) { std::fprintf(stderr, "%s\n", error.c_str()); return 1; }
  if (!silent) std::printf("%s -> %s\n", input.c_str(), output.c_str());
  return 0;
}
