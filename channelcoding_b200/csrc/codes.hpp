// codes.hpp -- host-side construction of the codes the engine decodes (header-only, C++17).
//
// Mirrors what the reference builds at start-up, cold path, never on the device:
//   GF(2^q) tables          math/galois.h:18-20 (default primitive polynomials), :269-301
//   primitive BCH g(x)      codes/bch.h:28-46, :62-78   (product over cyclotomic cosets of
//                           alpha^1, alpha^3, .., alpha^(2t-1))
//   RS g(x)                 codes/rs.h:18-28            (prod (x - alpha^(mu + i*step)))
//   h = (x^n + 1)/g, k = deg g, l = n - k, rate = l/n, dmin   codes/cyclic.h:270-280, :186-204
//   H(): k cyclic shifts of reversed h                        codes/cyclic.h:346-359
//   to_string "(n, l, dmin)-TAG"                              codes/cyclic.h:282-287
//   systematic encoding                                      codes/cyclic.h:35-40, :289-311
// Written from the mathematics, not from the reference's templates: polynomials are plain
// coefficient vectors over a table-driven field object.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace ccgpu {

struct Field {
  unsigned q = 0, size = 0, n = 0, poly = 0;
  std::vector<uint16_t> exp;  // 2*size entries: exp[p] = exp[p + n] = alpha^p, exp[2*size-1] = 0
  std::vector<uint16_t> log;  // size entries, log[0] = 0 by convention

  static unsigned default_poly(unsigned q) {
    static const unsigned table[9] = { 0, 0x3, 0x7, 0xb, 0x13, 0x25, 0x43, 0x83, 0x11d };
    if (q < 1 || q > 8) throw std::invalid_argument("GF(2^q): q must be in 1..8");
    return table[q];
  }
  Field() = default;
  explicit Field(unsigned q_, unsigned poly_ = 0) : q(q_), size(1u << q_), n((1u << q_) - 1), poly(poly_ ? poly_ : default_poly(q_)) {
    if (q < 1 || q > 8) throw std::invalid_argument("GF(2^q): q must be in 1..8");
    exp.assign(2 * size, 0);
    log.assign(size, 0);
    unsigned v = 1;
    for (unsigned p = 0; p < n; ++p) {
      exp[p] = exp[p + n] = static_cast<uint16_t>(v);
      log[v] = static_cast<uint16_t>(p);
      v <<= 1;
      if (v & size) v ^= poly;
    }
    exp[n] = 1;          // alpha^n = 1
    exp[2 * n] = 1;
    log[0] = 0;
  }
  unsigned mul(unsigned a, unsigned b) const { return (a && b) ? exp[log[a] + log[b]] : 0u; }
  unsigned div(unsigned a, unsigned b) const { return a ? exp[log[a] + n - log[b]] : 0u; }  // b != 0
  unsigned inv(unsigned a) const { return exp[n - log[a]]; }
  unsigned pow_alpha(unsigned p) const { return exp[p % n]; }
};

using Poly = std::vector<uint16_t>;  // low degree first

inline int degree(const Poly &p) {
  for (int i = static_cast<int>(p.size()) - 1; i >= 0; --i)
    if (p[i]) return i;
  return -1;
}
inline Poly poly_mul(const Field &F, const Poly &a, const Poly &b) {
  const int da = degree(a), db = degree(b);
  if (da < 0 || db < 0) return Poly{ 0 };
  Poly r(da + db + 1, 0);
  for (int i = 0; i <= da; ++i)
    if (a[i])
      for (int j = 0; j <= db; ++j) r[i + j] ^= static_cast<uint16_t>(F.mul(a[i], b[j]));
  return r;
}
// quotient and remainder; throws on a zero divisor
inline void poly_divmod(const Field &F, const Poly &a, const Poly &b, Poly *quo, Poly *rem) {
  const int da = degree(a), db = degree(b);
  if (db < 0) throw std::invalid_argument("polynomial division by zero");
  Poly r = a;
  Poly q(da >= db ? da - db + 1 : 1, 0);
  for (int d = da; d >= db; --d) {
    if (!r[d]) continue;
    const unsigned c = F.div(r[d], b[db]);
    q[d - db] = static_cast<uint16_t>(c);
    for (int j = 0; j <= db; ++j) r[d - db + j] ^= static_cast<uint16_t>(F.mul(b[j], c));
  }
  if (quo) *quo = q;
  if (rem) *rem = r;
}
inline unsigned poly_eval(const Field &F, const Poly &p, unsigned x) {
  unsigned acc = 0;
  for (int i = static_cast<int>(p.size()) - 1; i >= 0; --i) acc = F.mul(acc, x) ^ p[i];
  return acc;
}

struct CodeSpec {
  int family = 2;  // 0 BCH, 1 RS, 2 dense
  Field F;
  unsigned q = 0, n = 0, k = 0, l = 0, dmin = 0, t = 0, mu = 1, step = 1;
  double rate = 0.0;
  Poly g, h;
  std::vector<uint16_t> roots;  // syndrome evaluation points alpha^(..), 2t of them
  std::vector<uint8_t> H;       // rows x n, row-major
  unsigned rows = 0;

  // dmin as the reference reports it: (run of consecutive root exponents starting at 1) + 1,
  // over-counting by one when the run reaches the largest root exponent (cyclic.h:199-203).
  void finish() {
    Poly f(n + 1, 0);
    f[0] = 1;
    f[n] = 1;
    poly_divmod(F, f, g, &h, nullptr);
    h.resize(degree(h) + 1);
    k = static_cast<unsigned>(degree(g));
    l = n - k;
    rate = static_cast<double>(l) / n;
    std::vector<unsigned> expo;
    for (unsigned p = 0; p < n; ++p)
      if (poly_eval(F, g, F.exp[p]) == 0) expo.push_back(p);
    auto first = std::find(expo.begin(), expo.end(), 1u);
    unsigned run = 1;
    if (first != expo.end()) {
      auto it = first;
      while (it + 1 != expo.end() && *(it + 1) == *it + 1) ++it;
      run = static_cast<unsigned>(it - first) + 1;
      if (it + 1 == expo.end()) run += 1;  // no gap found: adjacent_find returns end()
    }
    dmin = run + 1;
    if (dmin > n) throw std::runtime_error("dmin > n");
    set_rows(k);
  }
  // parity-check matrix with `r` cyclic right-shifts of (h_l .. h_0 0 .. 0); r = k is H()
  void set_rows(unsigned r) {
    if (family == 2) throw std::invalid_argument("set_rows needs a BCH/RS code");
    if (r < 1 || r > n) throw std::invalid_argument("rows must be in 1..n");
    rows = r;
    H.assign(static_cast<size_t>(rows) * n, 0);
    for (unsigned i = 0; i < h.size(); ++i) H[i] = static_cast<uint8_t>(h[h.size() - 1 - i] ? 1 : 0);
    for (unsigned rr = 1; rr < rows; ++rr)
      for (unsigned c = 0; c < n; ++c) H[rr * n + (c + 1) % n] = H[(rr - 1) * n + c];
  }
  // "H from the roots of g(x)" -- cyclic::H_alt<T>() of codes/cyclic.h:361-385: row block i (i < t) holds the
  // binary expansion (q rows, least significant bit first) of alpha^(col * (2i + 1)).  as_reference = true
  // reproduces the reference's from_power (exponent reduced mod 2^q instead of mod n, galois.h:182-184,
  // SURVEY C4), which yields an invalid matrix for exponents >= 2^q; false reduces mod n.
  std::vector<uint8_t> h_alt(bool as_reference, unsigned *rows_out) const {
    if (family == 2) throw std::invalid_argument("h_alt needs a BCH/RS code");
    std::vector<uint8_t> M(static_cast<size_t>(t) * q * n, 0);
    for (unsigned i = 0; i < t; ++i)
      for (unsigned c = 0; c < n; ++c) {
        const unsigned power = c * (2 * i + 1);
        const unsigned v = as_reference ? F.exp[power % F.size] : F.exp[power % n];
        for (unsigned b = 0; b < q; ++b) M[(static_cast<size_t>(i) * q + b) * n + c] = (v >> b) & 1u;
      }
    *rows_out = t * q;
    return M;
  }
  std::string to_string(const std::string &tag) const {
    return "(" + std::to_string(n) + ", " + std::to_string(l) + ", " + std::to_string(dmin) + ")-" + tag;
  }
  // a(x) x^k + (a(x) x^k mod g(x))
  void encode(const uint8_t *msg, uint8_t *word) const {
    Poly xk(n, 0);
    for (unsigned i = 0; i < l; ++i) {
      if (msg[i] > F.n) throw std::invalid_argument("symbol is not a field element");
      xk[k + i] = msg[i];
    }
    Poly rem;
    poly_divmod(F, xk, g, nullptr, &rem);
    for (unsigned i = 0; i < n; ++i) word[i] = static_cast<uint8_t>(xk[i] ^ (i < k ? rem[i] : 0));
  }
};

// cap_kind 0: errors<v> -> t = v;  1: dmin<v> -> t = (v - 1) / 2      (codes/codes.h:15-26)
inline CodeSpec make_bch(unsigned q, int cap_kind, unsigned cap_value) {
  CodeSpec c;
  c.family = 0;
  c.F = Field(q);
  c.q = q;
  c.n = c.F.n;
  c.t = cap_kind == 1 ? (cap_value - 1) / 2 : cap_value;
  if (c.t < 1 || 2 * c.t >= c.n) throw std::invalid_argument("BCH: t out of range");
  std::vector<char> covered(c.n, 0);
  c.g = Poly{ 1 };
  for (unsigned i = 1; i < 2 * c.t; i += 2) {
    if (covered[i % c.n]) continue;  // same cyclotomic coset as an earlier root: lcm adds nothing
    Poly m{ 1 };
    unsigned e = i % c.n;
    do {
      covered[e] = 1;
      m = poly_mul(c.F, m, Poly{ static_cast<uint16_t>(c.F.exp[e]), 1 });
      e = (2 * e) % c.n;
    } while (e != i % c.n);
    c.g = poly_mul(c.F, c.g, m);
  }
  for (unsigned j = 1; j <= 2 * c.t; ++j) c.roots.push_back(c.F.exp[j % c.n]);
  c.finish();
  return c;
}

inline CodeSpec make_rs(unsigned q, unsigned t, unsigned mu, unsigned step) {
  CodeSpec c;
  c.family = 1;
  c.F = Field(q);
  c.q = q;
  c.n = c.F.n;
  c.t = t;
  c.mu = mu;
  c.step = step;
  if (t < 1 || 2 * t >= c.n) throw std::invalid_argument("RS: t out of range");
  c.g = Poly{ 1 };
  for (unsigned i = 0; i < 2 * t; ++i) {
    const uint16_t root = c.F.exp[(mu + i * step) % c.n];
    c.g = poly_mul(c.F, c.g, Poly{ root, 1 });
    c.roots.push_back(root);
  }
  c.finish();
  return c;
}

// structure of a dense parity-check matrix
struct HShape {
  int kind = 2;  // 0 cyclic no wrap, 1 cyclic with wrap, 2 general
  std::vector<int> taps;
  unsigned max_row_weight = 0, edges = 0;
};
inline HShape analyse_H(const uint8_t *H, unsigned rows, unsigned n) {
  HShape s;
  for (unsigned r = 0; r < rows; ++r) {
    unsigned w = 0;
    for (unsigned c = 0; c < n; ++c) w += H[r * n + c] ? 1 : 0;
    s.max_row_weight = std::max(s.max_row_weight, w);
    s.edges += w;
  }
  for (unsigned c = 0; c < n; ++c)
    if (H[c]) s.taps.push_back(static_cast<int>(c));
  bool cyclic = !s.taps.empty(), wrap = false;
  for (unsigned r = 1; r < rows && cyclic; ++r)
    for (unsigned c = 0; c < n; ++c)
      if ((H[r * n + (c + r) % n] != 0) != (H[c] != 0)) { cyclic = false; break; }
  if (cyclic) {
    wrap = static_cast<unsigned>(s.taps.back()) + rows - 1 >= n;
    s.kind = wrap ? 1 : 0;
  }
  return s;
}

}  // namespace ccgpu
