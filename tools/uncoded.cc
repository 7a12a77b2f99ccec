// tools/uncoded.cc -- the reference's `uncoded` program (src/simulation/uncoded.c++:1-55) on the GPU: the AWGN
// sweep of awgn_simulation over the `uncoded` pseudo-decoder of l symbols, same flags (--l, --seed), same log file
// "<l>-uncoded.log"; --out DIR, --max-samples N and --device D are additions like in tools/benchmark.cc.
#include <getopt.h>

#include <cstdlib>
#include <iostream>
#include <string>

#include "cc/simulation.h"

[[noreturn]] static void usage() {
  std::cout << "--l <num>          Choose code length l\n"
               "--seed <num>       Set seed of the random number generator.\n"
               "--out <dir>  --max-samples <num>  --device <num>" << std::endl;
  std::exit(EXIT_FAILURE);
}

int main(int argc, char *const argv[]) {
  unsigned l = 0;
  uint64_t seed = 0, cap = 1000000;
  std::string out = ".";
  int device = 0;
  static struct option options[] = {
    { "l", required_argument, nullptr, 'l' },           { "seed", required_argument, nullptr, 's' },
    { "out", required_argument, nullptr, 'o' },         { "max-samples", required_argument, nullptr, 'x' },
    { "device", required_argument, nullptr, 'd' },      { nullptr, 0, nullptr, 0 },
  };
  for (;;) {
    int idx = 0;
    const int c = getopt_long_only(argc, argv, "", options, &idx);
    if (c == -1) break;
    switch (c) {
    case 'l': l = static_cast<unsigned>(std::stoul(optarg)); break;
    case 's': seed = std::stoull(optarg); break;
    case 'o': out = optarg; break;
    case 'x': cap = std::stoull(optarg); break;
    case 'd': device = std::stoi(optarg); break;
    default: std::cerr << "Unkown argument: " << c << " " << std::endl; usage();
    }
  }
  if (!l) {
    std::cerr << "l not set." << std::endl;
    usage();
  }
  try {
    cc::decoder d = cc::uncoded(l, device);
    cc::awgn_simulation(d, 0.5, seed).samples_cap(cap).output_dir(out)();
  } catch (const std::exception &e) {
    std::cerr << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
