// oracle/integration_test.cc -- TEST INFRASTRUCTURE.  The reference's own classes (decoder,
// bitflip_simulation: src/simulation/simulation.{h,c++}) running on top of libccgpu.so through the
// adapter of INTEGRATION.md.  Built by oracle/build_ref.sh into oracle/_ref/integration_test.
//   integration_test <scratch dir>     (needs a GPU; writes the reference's "<name>.log" there)
#include <unistd.h>

#include <fstream>
#include <iostream>
#include <sstream>

#include "gpu_decoder.h"
#include "simulation/simulation.h"

int main(int argc, char **argv) {
  if (argc < 2 || chdir(argv[1]) != 0) return 2;
  // the reference's type-erased decoder holding the GPU adapter (simulation.h:58-60)
  decoder ms = gpu_decoder<cyclic::primitive_bch<5, dmin<7> >, CCGPU_MS>();
  decoder nms = gpu_decoder<cyclic::primitive_bch<6, errors<5> >, CCGPU_NMS>(0.8);
  if (ms.to_string() != "(31, 16, 7)-MS" || nms.to_string() != "(63, 36, 11)-NMS" || nms.n() != 63) return 3;  // code part from the reference object, tag from the variant
  // the reference's own exhaustive bit-flip loop (simulation.c++:156-213) calling correct() per pattern
  bitflip_simulation(ms, 3)();
  std::ifstream f("(31, 16, 7)-MS.log");
  std::stringstream ss;
  ss << f.rdbuf();
  std::cout << ss.str();
  // the reference's AWGN inner loop (simulation.c++:124-136) on a few frames
  std::mt19937_64 generator(0);
  std::normal_distribution<float> distribution(1.0, 0.75f);
  std::vector<float> b(nms.n());
  size_t word_errors = 0, frames = 2000;
  for (size_t i = 0; i < frames; i++) {
    std::generate(std::begin(b), std::end(b), [&] { return distribution(generator); });
    try {
      auto result = nms.correct(b);
      if (std::any_of(std::cbegin(result), std::cend(result), [](const auto &bit) { return bool(bit); })) word_errors++;
    } catch (const decoding_failure &) {
      word_errors++;
    }
  }
  std::cout << "awgn " << word_errors << " " << frames << std::endl;
  return 0;
}
