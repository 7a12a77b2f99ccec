// tools/bitflips.cc -- the reference's `bitflips` program (src/simulation/bitflips.c++:1-40, the source of Table 3 of
// its report) on the GPU: the nine decoders of BCH(31,16,7) under bitflip_simulation(decoder, 6), i.e. every
// pattern of 0..6 flipped bits (736 281 words of weight 6 alone), one "<name>.log" per decoder.
// Additions: --out DIR, --errors W (default 6 like the reference), --device D.
#include <getopt.h>

#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "cc/simulation.h"

int main(int argc, char *const argv[]) {
  std::string out = ".";
  unsigned errors = 6;
  int device = 0;
  static struct option options[] = {
    { "out", required_argument, nullptr, 'o' },
    { "errors", required_argument, nullptr, 'e' },
    { "device", required_argument, nullptr, 'd' },
    { nullptr, 0, nullptr, 0 },
  };
  for (;;) {
    int idx = 0;
    const int c = getopt_long_only(argc, argv, "", options, &idx);
    if (c == -1) break;
    if (c == 'o') out = optarg;
    else if (c == 'e') errors = static_cast<unsigned>(std::stoul(optarg));
    else if (c == 'd') device = std::stoi(optarg);
    else return EXIT_FAILURE;
  }
  try {
    using namespace cc;
    const std::vector<decoder> decoders{
      primitive_bch<5, dmin<7>, berlekamp_massey_tag>(device),
      primitive_bch<5, dmin<7>, peterson_gorenstein_zierler_tag>(device),
      primitive_bch<5, dmin<7>, euklid_tag>(device),
      primitive_bch<5, dmin<7>, min_sum_tag<50> >(device),
      primitive_bch<5, dmin<7>, normalized_min_sum_tag<50, std::ratio<8, 10> > >(device),
      primitive_bch<5, dmin<7>, offset_min_sum_tag<50, std::ratio<1, 100> > >(device),
      primitive_bch<5, dmin<7>, self_correcting_1_min_sum_tag<50> >(device),
      primitive_bch<5, dmin<7>, self_correcting_2_min_sum_tag<50> >(device),
      primitive_bch<5, dmin<7>, normalized_2d_min_sum_tag<50> >(device),
    };
    for (const auto &d : decoders) std::cout << d.to_string() << std::endl;
    thread_pool p;
    for (const auto &d : decoders) p.push(bitflip_simulation(d, errors).output_dir(out));
  } catch (const std::exception &e) {
    std::cerr << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
