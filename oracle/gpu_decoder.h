// oracle/gpu_decoder.h -- the adapter of INTEGRATION.md (section A), verbatim, so that it is compiled
// against the REAL reference headers by oracle/build_ref.sh (-> oracle/_ref/integration_test) and run
// on the GPU box by tests/test_integration.py: the reference's own decoder / bitflip_simulation /
// awgn loop drive libccgpu.so through it.  Test infrastructure; the product does not use this file.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

// src/codes/gpu_decoder.h  (new file in the reference)
#include <ccgpu.h>
#include "codes/bch.h"

template <typename Code, int Variant /* CCGPU_MS .. CCGPU_OMS_Q */>
class gpu_decoder {
  Code code;                                    // the reference's own object: g, h, H(), rate, to_string
  std::shared_ptr<ccgpu_ctx> ctx;
  std::shared_ptr<ccgpu_code> dev;
  ccgpu_ms_params params;
public:
  static constexpr unsigned n = Code::n;
  const double rate;
  gpu_decoder(double alpha = 1.0, double beta = 0.0, unsigned iterations = 50) : rate(code.rate) {
    ccgpu_ctx *c = nullptr;
    if (ccgpu_create(0, &c) != CCGPU_OK) throw std::runtime_error("no CUDA device");
    ctx.reset(c, ccgpu_destroy);
    const auto H = code.template H<uint8_t>();   // cyclic.h:346-359, built ONCE instead of per frame (:265)
    std::vector<uint8_t> flat;
    for (size_t r = 0; r < H.rows(); r++) flat.insert(flat.end(), H.at(r).begin(), H.at(r).end());
    ccgpu_code *d = nullptr;
    if (ccgpu_code_from_dense(c, flat.data(), H.rows(), H.columns(), code.rate, &d) != CCGPU_OK)
      throw std::runtime_error(ccgpu_last_error(c));
    dev.reset(d, ccgpu_code_destroy);
    params = ccgpu_ms_params{ Variant, CCGPU_STOP_REF_ZERO_OVERLAP, iterations, 0, alpha, beta };
  }
  // "(n, l, dmin)-TAG" (cyclic.h:282-287) with the tag of the VARIANT this decoder runs, not of the Code's algebraic
  // default: the string is the log-file name (simulation.c++:98), two variants of one code must not collide
  std::string to_string() const {
    static const char *const tags[] = { "MS", "NMS", "OMS", "SCMS1", "SCMS2", "2DNMS", "SPA", "MSQ", "NMSQ", "OMSQ" };
    const std::string s = code.to_string();
    return s.substr(0, s.find_last_of('-') + 1) + tags[Variant];
  }
  template <typename R> std::vector<R> correct(const std::vector<float> &b) const {
    std::vector<uint8_t> bits(n);
    uint8_t failed = 0;
    if (ccgpu_decode_llr(ctx.get(), dev.get(), &params, b.data(), 1, bits.data(), nullptr, nullptr, &failed))
      throw std::runtime_error(ccgpu_last_error(ctx.get()));
    if (failed) throw decoding_failure("Decoding failure");          // soft_decision.h:201
    return std::vector<R>(bits.begin(), bits.end());                  // R = math::ef_element<2,1> or uint8_t
  }
  // the batched call the simulation should use instead of the per-frame loop (simulation.c++:124-136)
  ccgpu_counters awgn_point(double ebno_db, uint64_t frames, uint64_t seed, uint32_t point) const {
    ccgpu_counters c{};
    ccgpu_awgn_point(ctx.get(), dev.get(), &params, ebno_db, seed, point, 0, frames, &c);
    return c;
  }
};
// benchmark.c++:28  ->  decoders{ gpu_decoder<cyclic::primitive_bch<6, errors<5>>, CCGPU_NMS>(0.8), ... }
