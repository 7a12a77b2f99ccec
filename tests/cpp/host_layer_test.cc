// tests/cpp/host_layer_test.cc -- exercises the C++ drop-in layer (include/cc/*.h) the way the
// reference's own programs use their classes (exercises.c++, bitflips.c++, benchmark.c++).
//   host_layer_test --host        no GPU needed: tags, capabilities, sweep start rule
//   host_layer_test --gpu <dir>   needs a B200: correct(), decoding_failure, to_string, bit-flip KAT,
//                                 awgn sweep log format
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <cstdlib>

#include "cc/simulation.h"

using namespace cc;

#define CHECK(cond)                                                                  \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      std::cerr << "CHECK failed: " #cond " at line " << __LINE__ << std::endl;      \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

static int host_tests() {
  // codes/codes.h:15-26
  static_assert(correction_capability<dmin<7> >::value == 3, "");
  static_assert(correction_capability<dmin<6> >::value == 2, "");
  static_assert(correction_capability<errors<5> >::value == 5, "");
  // tag names (soft_decision.h:20-73, hard_decision.h:15-24)
  CHECK(min_sum_tag<50>::to_string() == "MS");
  CHECK((normalized_min_sum_tag<50, std::ratio<8, 10> >::to_string() == "NMS"));
  CHECK((offset_min_sum_tag<50, std::ratio<1, 100> >::to_string() == "OMS"));
  CHECK(self_correcting_1_min_sum_tag<50>::to_string() == "SCMS1");
  CHECK(self_correcting_2_min_sum_tag<50>::to_string() == "SCMS2");
  CHECK(normalized_2d_min_sum_tag<50>::to_string() == "2DNMS");
  CHECK(berlekamp_massey_tag::to_string() == "BM" && euklid_tag::to_string() == "EUKLID" &&
        peterson_gorenstein_zierler_tag::to_string() == "PGZ");
  // extension tags: fixed-point min-sum
  CHECK((fixed_normalized_min_sum_tag<50, std::ratio<8, 10> >::to_string() == "NMSQ") && fixed_min_sum_tag<>::to_string() == "MSQ");
  {
    const ccgpu_ms_params p = detail::params_of<fixed_offset_min_sum_tag<20, std::ratio<1, 4>, fixed_point<16, 63, 40> > >::get(stop_rule::gf2_parity);
    CHECK(p.variant == CCGPU_OMS_Q && p.max_iter == 20 && p.beta == 0.25 && p.q_scale == 16.0 && p.q_y_max == 63 && p.q_msg_max == 40 &&
          p.stop_rule == CCGPU_STOP_GF2_PARITY);
    const ccgpu_ms_params f = detail::params_of<min_sum_tag<50> >::get(stop_rule::reference);
    CHECK(f.q_scale == 0.0 && f.q_y_max == 0 && f.q_msg_max == 0);
  }
  CHECK((normalized_min_sum_tag<50, std::ratio<8, 10> >::alpha == 0.8));
  CHECK((offset_min_sum_tag<50, std::ratio<1, 100> >::beta == 0.01));
  // the reference's 2DNMS default is alpha = beta = 1 (soft_decision.h:71)
  CHECK(normalized_2d_min_sum_tag<50>::alpha == 1.0 && normalized_2d_min_sum_tag<50>::beta == 1.0);
  CHECK((normalized_2d_min_sum_tag<50, std::ratio<968, 1000>, std::ratio<907, 1000> >::beta == 907.0 / 125.0));
  // sweep start, simulation.c++:105-107 with the Shannon-limit table :21-52 (values from SURVEY 8d)
  CHECK(awgn_simulation::start_ebno(36.0 / 63, 0.5) == 1.5);
  CHECK(awgn_simulation::start_ebno(7.0 / 15, 0.5) == 1.0);
  CHECK(awgn_simulation::start_ebno(64.0 / 127, 0.5) == 1.0);
  CHECK(awgn_simulation::start_ebno(131.0 / 255, 0.5) == 1.0);
  CHECK(ccgpu_shannon_limit_db(0.495) == 0.188);   // table: rates (0.49, 0.50] -> 0.188 dB
  CHECK(ccgpu_shannon_limit_db(0.795) == 2.045);   // table: rate 0.80 -> 2.045 dB
  CHECK(ccgpu_shannon_limit_db(0.005) == -1.548);  // table: rate 0.01 -> -1.548 dB
  CHECK(ccgpu_shannon_limit_db(26.0 / 31) == 2.503 && awgn_simulation::start_ebno(26.0 / 31, 0.5) == 3.5);  // BCH(31,26)
  CHECK(std::fabs(ccgpu_shannon_limit_db_numeric(0.5) - 0.188) < 0.003);
  std::cout << "host tests ok" << std::endl;
  return 0;
}

static int gpu_tests(const std::string &dir) {
  // exercises.c++ task 6.1: primitive_bch<4, dmin<7>> corrects b1 and b2 to a
  {
    primitive_bch<4, dmin<7> > code;
    const std::vector<unsigned char> a({ 1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 1, 0, 1, 0, 1 });
    std::vector<unsigned> b1({ 1, 1, 1, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 1, 1 });
    std::vector<unsigned> b2({ 1, 1, 1, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 1 });
    CHECK(code.correct(b1) == a);
    CHECK(code.correct(b2) == a);
    CHECK(code.to_string() == "(15, 5, 7)-PGZ");
    CHECK(decltype(code)::n == 15 && decltype(code)::t == 3);
  }
  // task 6.2: decoding failure is an exception of type decoding_failure
  {
    primitive_bch<4, dmin<5> > code;
    const std::vector<unsigned> b({ 1, 0, 0, 1, 0, 1, 1, 1, 1, 0, 1, 1, 0, 0, 0 });
    bool threw = false;
    try {
      code.correct(b);
    } catch (const decoding_failure &) {
      threw = true;
    }
    CHECK(threw);
  }
  // task 6.10: three algorithm tags, one answer
  {
    const std::vector<uint8_t> a({ 1, 0, 1, 0, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1 });
    const std::vector<uint8_t> want({ 1, 0, 1, 0, 0, 1, 1, 1, 1, 0, 1, 0, 1, 0, 1 });
    CHECK((primitive_bch<4, errors<2>, peterson_gorenstein_zierler_tag>().correct(a) == want));
    CHECK((primitive_bch<4, errors<2>, berlekamp_massey_tag>().correct(a) == want));
    CHECK((primitive_bch<4, errors<2>, euklid_tag>().correct(a) == want));
  }
  // RS(255,223) through the same interface
  {
    rs<8, errors<16>, euklid_tag> code;
    CHECK(code.to_string() == "(255, 223, 34)-EUKLID");  // sic: dmin as the reference computes it
    std::vector<uint8_t> w(255, 0);
    w[3] = 7; w[100] = 200; w[254] = 1;
    CHECK(code.correct(w) == std::vector<uint8_t>(255, 0));
  }
  // exercises.c++ task 6.7 / 6.8: RS(7,3) with erasures through correct(b, erasures)
  {
    rs<3, errors<2>, berlekamp_massey_tag> code;
    // alpha^p table of GF(8), primitive polynomial x^3 + x + 1: 1 2 4 3 6 7 5
    const uint8_t a[7] = { 1, 2, 4, 3, 6, 7, 5 };
    const std::vector<uint8_t> word({ a[6], a[2], a[2], a[5], a[4], a[6], a[5] });
    std::vector<uint8_t> b(word);
    const std::vector<unsigned> erasures({ 5, 4, 3, 2 });
    for (unsigned e : erasures) b[e] = 0;
    CHECK(code.correct(b, erasures) == word);
    const std::vector<uint8_t> b8({ a[2], a[0], a[4], a[0], a[5], a[0], a[2] });
    const std::vector<uint8_t> want8({ a[2], a[5], a[4], a[6], a[5], a[6], a[2] });
    CHECK(code.correct(b8, std::vector<unsigned>({ 1, 3 })) == want8);
  }
  // soft decoding with erasures: the erased positions get channel value 0 (cyclic.h:261-262)
  {
    primitive_bch<6, errors<5>, normalized_min_sum_tag<50, std::ratio<8, 10> > > code;
    std::vector<float> y(63, 1.0f);
    y[7] = -3.0f;
    y[20] = -2.0f;
    CHECK(code.correct(y, std::vector<unsigned>({ 7, 20 })) == std::vector<uint8_t>(63, 0));
  }
  // soft decoders through the type-erased decoder, like simulation.c++:124-136
  {
    decoder d = primitive_bch<6, errors<5>, normalized_min_sum_tag<50, std::ratio<8, 10> > >();
    CHECK(d.to_string() == "(63, 36, 11)-NMS" && d.n() == 63 && std::fabs(d.rate() - 36.0 / 63) < 1e-15);
    std::vector<float> y(63, 1.0f);
    y[5] = -0.4f;
    const auto r = d.correct(y);
    CHECK(r == std::vector<uint8_t>(63, 0));
    std::vector<float> bad(63, -1.0f);  // the all-one word never satisfies the reference's stop rule
    bool threw = false;
    try {
      d.correct(bad);
    } catch (const decoding_failure &) {
      threw = true;
    }
    CHECK(threw);
    bool size_error = false;
    try {
      d.correct(std::vector<float>(10, 1.0f));
    } catch (const std::runtime_error &) {
      size_error = true;
    }
    CHECK(size_error);
  }
  // bitflips.c++ on (31,16,7): failures per weight and the log format
  {
    decoder ms = primitive_bch<5, dmin<7>, min_sum_tag<50> >();
    decoder bm = primitive_bch<5, dmin<7>, berlekamp_massey_tag>();
    CHECK(ms.bitflip_point(2).frame_errors == 138 && ms.bitflip_point(3).frame_errors == 3557);
    CHECK(bm.bitflip_point(3).frame_errors == 0 && bm.bitflip_point(4).frame_errors == 31465);
    bitflip_simulation(ms, 3).output_dir(dir)();
    std::ifstream f(dir + "/(31, 16, 7)-MS.log");
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string expect = " errors                   wer\n"
                               "      0 0.000000000000000e+00\n"
                               "      1 0.000000000000000e+00\n"
                               "      2 2.967741935483871e-01\n"
                               "      3 7.913236929922136e-01\n";
    if (ss.str() != expect) std::cerr << "got:\n" << ss.str() << "want:\n" << expect;
    CHECK(ss.str() == expect);
    bool refused = false;  // simulation.c++:72-81: an existing log is never overwritten
    try {
      bitflip_simulation(ms, 1).output_dir(dir)();
    } catch (const std::runtime_error &) {
      refused = true;
    }
    CHECK(refused);
  }
  // awgn sweep: schedule and log format (simulation.c++:95-150)
  {
    decoder d = primitive_bch<4, errors<2>, min_sum_tag<50> >();
    awgn_simulation(d, 0.5, 0).samples_cap(200000).output_dir(dir)();
    std::ifstream f(dir + "/(15, 7, 5)-MS.log");
    std::string line;
    std::getline(f, line);
    CHECK(line == "   ebno                   wer");
    int points = 0;
    double first = -1, last_wer = 1.0;
    while (std::getline(f, line)) {
      double eb, wer;
      std::istringstream(line) >> eb >> wer;
      if (points == 0) first = eb;
      CHECK(wer <= last_wer * 1.5 + 1e-3);  // a waterfall
      last_wer = wer;
      ++points;
    }
    CHECK(first == 1.0 && points == 15);  // 1.0, 1.5, .., 8.0
    CHECK(last_wer < 1e-3);
  }
  // extension: multiple bases through the C++ layer -- one rotation 0 equals correct_batch, more rotations decode
  // at least the frames the first one decodes
  {
    primitive_bch<6, errors<5>, normalized_min_sum_tag<50, std::ratio<8, 10> > > code;
    const unsigned n = 63, frames = 3000;
    std::vector<float> y(size_t(frames) * n);
    uint64_t lcg = 12345;
    for (auto &v : y) {  // crude noise, enough to make a good share of the frames fail on one matrix
      float acc = 0;
      for (int k = 0; k < 12; ++k) {
        lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
        acc += float((lcg >> 40) & 0xffff) / 65536.0f;
      }
      v = 1.0f + 0.66f * (acc - 6.0f);
    }
    std::vector<uint8_t> b1(y.size()), f1(frames), b2(y.size()), f2(frames), b3(y.size()), f3(frames), chosen(frames);
    code.correct_batch(y.data(), frames, b1.data(), f1.data());
    code.correct_batch_multiple_bases(y.data(), frames, { 0u }, b2.data(), f2.data());
    CHECK(b1 == b2 && f1 == f2);
    code.correct_batch_multiple_bases(y.data(), frames, { 0u, 9u, 20u, 33u, 47u }, b3.data(), f3.data(), nullptr, nullptr,
                                      chosen.data());
    unsigned fail1 = 0, fail3 = 0, moved = 0;
    for (unsigned f = 0; f < frames; ++f) {
      fail1 += f1[f];
      fail3 += f3[f];
      moved += chosen[f] != 0;
      if (!f1[f]) CHECK(!f3[f]);
    }
    CHECK(fail1 > 50 && fail3 < fail1 && moved > 0);
  }
  std::cout << "gpu tests ok" << std::endl;
  return 0;
}

static bool same(const ccgpu_counters &a, const ccgpu_counters &b) {
  return a.frames == b.frames && a.frame_errors == b.frame_errors && a.bit_errors == b.bit_errors &&
         a.iterations == b.iterations && a.failures == b.failures && a.undetected == b.undetected;
}

// device groups: the same objects created under a cc::device_group shard every point over the members and must
// return the counters of the one-device run (the noise is keyed by the global frame index).  `devices` may name
// one device several times (then the sharding / merging is exercised on a single GPU).
static int group_tests(const std::string &dir, const std::vector<int> &devices) {
  using nms = normalized_min_sum_tag<50, std::ratio<8, 10> >;
  using nmsq = fixed_normalized_min_sum_tag<50, std::ratio<8, 10> >;
  device_group::use(std::vector<int>());
  decoder s_nms = primitive_bch<6, errors<5>, nms>(), s_q = primitive_bch<6, errors<5>, nmsq>(),
          s_bm = primitive_bch<6, errors<5>, berlekamp_massey_tag>(), s_big = primitive_bch<8, errors<18>, nms>(),
          s_ms31 = primitive_bch<5, dmin<7>, min_sum_tag<50> >(), s_un = uncoded(100);
  device_group::use(devices);
  CHECK(device_group::current() && device_group::current()->size() == static_cast<int>(devices.size()));
  decoder g_nms = primitive_bch<6, errors<5>, nms>(), g_q = primitive_bch<6, errors<5>, nmsq>(),
          g_bm = primitive_bch<6, errors<5>, berlekamp_massey_tag>(), g_big = primitive_bch<8, errors<18>, nms>(),
          g_ms31 = primitive_bch<5, dmin<7>, min_sum_tag<50> >(), g_un = uncoded(100);
  CHECK(g_q.to_string() == "(63, 36, 11)-NMSQ");
  const uint64_t sizes[] = { 1, 7, 16384, 100003, 1000000 };
  for (uint64_t frames : sizes) {
    CHECK(same(s_nms.awgn_point(4.0, frames, 3, 5, 77), g_nms.awgn_point(4.0, frames, 3, 5, 77)));
    CHECK(same(s_q.awgn_point(4.0, frames, 3, 5, 77), g_q.awgn_point(4.0, frames, 3, 5, 77)));
    CHECK(same(s_bm.awgn_point(5.0, frames, 3, 5, 77), g_bm.awgn_point(5.0, frames, 3, 5, 77)));
    CHECK(same(s_un.awgn_point(5.0, frames, 3, 5, 77), g_un.awgn_point(5.0, frames, 3, 5, 77)));
  }
  CHECK(same(s_big.awgn_point(6.0, 50001, 1, 2), g_big.awgn_point(6.0, 50001, 1, 2)));
  for (unsigned w = 0; w <= 4; ++w) CHECK(same(s_ms31.bitflip_point(w), g_ms31.bitflip_point(w)));
  // batched decode with host buffers, sharded: identical outputs
  {
    device_group::use(std::vector<int>());
    primitive_bch<6, errors<5>, nms> one;
    device_group::use(devices);
    primitive_bch<6, errors<5>, nms> many;
    const unsigned n = 63, frames = 70001;
    std::vector<float> y(size_t(frames) * n);
    uint64_t lcg = 99;
    for (auto &v : y) {
      lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
      v = 1.0f + 1.6f * (float((lcg >> 40) & 0xffff) / 65536.0f - 0.5f) * 1.7f;
    }
    std::vector<uint8_t> b1(y.size()), f1(frames), i1(frames), b2(y.size()), f2(frames), i2(frames);
    one.correct_batch(y.data(), frames, b1.data(), f1.data(), i1.data());
    many.correct_batch(y.data(), frames, b2.data(), f2.data(), i2.data());
    CHECK(b1 == b2 && f1 == f2 && i1 == i2);
  }
  // the whole sweep: byte-identical log
  {
    const std::string d1 = dir + "/one", dn = dir + "/group";
    CHECK(::mkdir(d1.c_str(), 0755) == 0 && ::mkdir(dn.c_str(), 0755) == 0);
    awgn_simulation(s_nms, 0.5, 0).samples_cap(300000).output_dir(d1)();
    awgn_simulation(g_nms, 0.5, 0).samples_cap(300000).output_dir(dn)();
    std::ifstream a(d1 + "/(63, 36, 11)-NMS.log"), b(dn + "/(63, 36, 11)-NMS.log");
    std::stringstream sa, sb;
    sa << a.rdbuf();
    sb << b.rdbuf();
    CHECK(!sa.str().empty() && sa.str() == sb.str());
  }
  device_group::use(std::vector<int>());
  std::cout << "group tests ok (" << devices.size() << " members)" << std::endl;
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 2 && std::string(argv[1]) == "--host") return host_tests();
  if (argc >= 4 && std::string(argv[1]) == "--group") {  // --group <scratch dir> <device> [<device> ...]
    std::vector<int> devices;
    for (int i = 3; i < argc; ++i) devices.push_back(std::atoi(argv[i]));
    return group_tests(argv[2], devices);
  }
  if (argc >= 3 && std::string(argv[1]) == "--gpu") return gpu_tests(argv[2]);
  std::cerr << "usage: host_layer_test --host | --gpu <scratch dir>" << std::endl;
  return 2;
}
