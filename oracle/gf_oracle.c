/* oracle/gf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's GF(2^q) arithmetic, code construction and
 * hard-decision (algebraic) decoder (hannesweisbach/channelcoding):
 *   tables            src/math/galois.h:269-301  (init_tables)
 *   poly arithmetic   src/math/polynomial.h:41-75 (division), :207-231 (product), :273-284 (Horner)
 *   BCH g(x)          src/codes/bch.h:28-46, :62-78      RS g(x)  src/codes/rs.h:18-28
 *   h, k, l, dmin     src/codes/cyclic.h:270-280, :186-204      H()  src/codes/cyclic.h:346-359
 *   encode            src/codes/cyclic.h:35-40, :289-311
 *   correct (hard)    src/codes/cyclic.h:207-252 with the Euklid/Sugiyama locator of
 *                     src/codes/hard_decision.h:157-196, brute-force roots (polynomial.h:16-28,
 *                     cyclic.h:126-150), error values 1 (bch.h:80-83) or the naive linear system
 *                     (rs.h:41-78, linear_equation_system.h) and the re-syndrome check.
 * The reference's Berlekamp-Massey reads out of bounds (hard_decision.h:137-141),
 * so Euklid is the algebraic oracle.  Pinned against oracle/_ref/libccref.so
 * (tests/test_oracle_pin.py) and the exercises.c++ known answers (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXP 1024

typedef struct {
  unsigned q, size, n;
  uint16_t exp[2 * 256 * 2]; /* q <= 9 head-room; reference supports q <= 8 by default */
  uint16_t log[512];
} gf_t;

/* galois.h:18-20 default primitive polynomials */
static const unsigned default_poly[9] = { 0, 0x3, 0x7, 0xb, 0x13, 0x25, 0x43, 0x83, 0x11d };

/* galois.h:269-301.  exp has 2*size entries: exp[p] = exp[p + size - 1] = alpha^p;
 * exp[2*size-1] stays 0 (value-initialised). */
int oracle_gf_init(gf_t *gf, unsigned q, unsigned poly) {
  if (q < 1 || q > 8) return -1;
  if (!poly) poly = default_poly[q];
  memset(gf, 0, sizeof(*gf));
  gf->q = q;
  gf->size = 1u << q;
  gf->n = gf->size - 1;
  unsigned polynomial = 1;
  for (unsigned power = 0; power < gf->size - 1; power++) {
    gf->log[polynomial] = (uint16_t)power;
    gf->exp[power] = (uint16_t)polynomial;
    gf->exp[power + gf->size - 1] = (uint16_t)polynomial;
    unsigned carry = polynomial & (1u << (q - 1));
    polynomial <<= 1;
    if (carry) polynomial ^= poly;
    polynomial &= 0xffffu;
  }
  gf->log[0] = 0;
  gf->exp[gf->size - 1] = 1;
  gf->exp[2 * gf->size - 2] = 1;
  return 0;
}

/* flat copies for the python side */
int oracle_gf_tables(unsigned q, unsigned poly, uint16_t *exp_out, uint16_t *log_out) {
  gf_t gf;
  if (oracle_gf_init(&gf, q, poly)) return -1;
  memcpy(exp_out, gf.exp, 2 * gf.size * sizeof(uint16_t));
  memcpy(log_out, gf.log, gf.size * sizeof(uint16_t));
  return 0;
}

/* galois.h:194-210 */
static unsigned gmul(const gf_t *gf, unsigned a, unsigned b) {
  if (!a || !b) return 0;
  return gf->exp[gf->log[a] + gf->log[b]];
}
static unsigned gdiv(const gf_t *gf, unsigned a, unsigned b) { /* b != 0 */
  if (!a) return 0;
  return gf->exp[gf->log[a] - gf->log[b] + gf->size - 1];
}
static unsigned ginv(const gf_t *gf, unsigned a) { return gdiv(gf, 1, a); }
/* ---- polynomials: coefficient arrays, low degree first (polynomial.h:31-330) ------------- */
typedef struct {
  int len;
  uint16_t c[MAXP];
} poly_t;

static int pdeg(const poly_t *p) { /* polynomial.h:131-135; -1 for the zero polynomial */
  for (int i = p->len - 1; i >= 0; i--)
    if (p->c[i]) return i;
  return -1;
}
static void pset1(poly_t *p, unsigned v) { p->len = 1; p->c[0] = (uint16_t)v; }
static void padd(poly_t *a, const poly_t *b) { /* a += b */
  while (a->len < b->len) a->c[a->len++] = 0;
  for (int i = 0; i < b->len; i++) a->c[i] ^= b->c[i];
}
static void pmul(const gf_t *gf, poly_t *out, const poly_t *a, const poly_t *b) {
  int da = pdeg(a), db = pdeg(b);
  poly_t r;
  if (da < 0 || db < 0) { pset1(out, 0); return; } /* polynomial.h:208-209 */
  r.len = da + db + 1;
  memset(r.c, 0, sizeof(uint16_t) * r.len);
  for (int i = 0; i <= da; i++)
    if (a->c[i])
      for (int j = 0; j <= db; j++) r.c[i + j] ^= (uint16_t)gmul(gf, a->c[i], b->c[j]);
  *out = r;
}
/* polynomial.h:41-75; returns -1 on division by the zero polynomial (std::logic_error) */
static int pdivmod(const gf_t *gf, poly_t *quo, poly_t *rem, const poly_t *a, const poly_t *b) {
  int da = pdeg(a), db = pdeg(b);
  if (db < 0) return -1;
  poly_t q, r = *a;
  if (da < db) {
    pset1(&q, 0);
  } else {
    q.len = da - db + 1;
    memset(q.c, 0, sizeof(uint16_t) * q.len);
    for (int d = da; d >= db; d--) {
      unsigned coef = gdiv(gf, r.c[d], b->c[db]);
      if (!coef) continue;
      q.c[d - db] ^= (uint16_t)coef;
      for (int j = 0; j <= db; j++) r.c[d - db + j] ^= (uint16_t)gmul(gf, b->c[j], coef);
    }
  }
  if (quo) *quo = q;
  if (rem) *rem = r;
  return 0;
}
/* polynomial.h:273-284: Horner from the highest stored coefficient; 0 for x == 0 */
static unsigned peval(const gf_t *gf, const poly_t *p, unsigned x) {
  if (p->len == 0 || x == 0) return 0;
  unsigned result = p->c[p->len - 1];
  for (int i = p->len - 2; i >= 0; i--) result = gmul(gf, result, x) ^ p->c[i];
  return result;
}

/* ---- code construction -------------------------------------------------------------------- */
typedef struct {
  gf_t gf;
  int family; /* 0 BCH, 1 RS */
  unsigned n, t, k, l, dmin, nroots;
  double rate;
  poly_t g, h;
  uint16_t roots[512]; /* syndrome evaluation points (bch.h:48-55 / rs.h:30-39) */
} code_t;

static void pgcd(const gf_t *gf, poly_t *out, const poly_t *a, const poly_t *b) {
  /* polynomial.h:289-294 */
  poly_t x = *a, y = *b, r;
  while (pdeg(&y) >= 0) {
    pdivmod(gf, 0, &r, &x, &y);
    x = y;
    y = r;
  }
  *out = x;
}

/* cyclic.h:186-204 consecutive_zeroes(g) */
static unsigned consecutive_zeroes(const gf_t *gf, const poly_t *g) {
  unsigned powers[512], np = 0;
  for (unsigned p = 1; p <= gf->n; p++) { /* field iteration order alpha^1 .. alpha^(n-1), 1 */
    unsigned el = gf->exp[p % gf->n];
    if (peval(gf, g, el) == 0) powers[np++] = gf->log[el];
  }
  for (unsigned i = 1; i < np; i++) /* std::sort */
    for (unsigned j = i; j > 0 && powers[j - 1] > powers[j]; j--) {
      unsigned tmp = powers[j]; powers[j] = powers[j - 1]; powers[j - 1] = tmp;
    }
  unsigned first = 0;
  while (first < np && powers[first] != 1) first++; /* std::find(.., 1) */
  unsigned last = np;                               /* std::adjacent_find(first, end, lhs+1 != rhs) */
  for (unsigned i = first; i + 1 < np; i++)
    if (powers[i] + 1 != powers[i + 1]) { last = i; break; }
  if (first >= np) last = np;
  return (last - first) + 1;
}

static int finish_code(code_t *c) {
  /* cyclic.h:270-280 */
  poly_t f;
  f.len = (int)c->n + 1;
  memset(f.c, 0, sizeof(uint16_t) * f.len);
  f.c[0] = 1;
  f.c[c->n] = 1; /* x^n + 1 (cyclic.h:120-123) */
  pdivmod(&c->gf, &c->h, 0, &f, &c->g);
  c->k = (unsigned)pdeg(&c->g);
  c->l = c->n - c->k;
  c->dmin = consecutive_zeroes(&c->gf, &c->g) + 1;
  c->rate = (double)c->l / c->n;
  return c->dmin > c->n ? -1 : 0;
}

/* bch.h:28-46 with minimal_polynomial_roots :62-78; syndromes() :48-55 */
int oracle_bch_init(code_t *c, unsigned q, unsigned t) {
  memset(c, 0, sizeof(*c));
  if (oracle_gf_init(&c->gf, q, 0)) return -1;
  c->family = 0;
  c->n = c->gf.n;
  c->t = t;
  pset1(&c->g, 1);
  for (unsigned p = 1; p < 2 * t; p += 2) {
    poly_t m, lin, gc, quo, prod;
    pset1(&m, 1);
    unsigned root = p;
    for (unsigned r = 0; r < q; r++) {
      if (r > 0) {
        root = ((1u << r) * p) % c->n;
        if (root == p) break;
      }
      lin.len = 2;
      lin.c[0] = (uint16_t)c->gf.exp[root % c->gf.size]; /* from_power as written (galois.h:182) */
      lin.c[1] = 1;
      pmul(&c->gf, &prod, &m, &lin);
      m = prod;
    }
    /* g = g.lcm(m) = g / gcd(g, m) * m   (polynomial.h:296-301) */
    pgcd(&c->gf, &gc, &c->g, &m);
    pdivmod(&c->gf, &quo, 0, &c->g, &gc);
    pmul(&c->gf, &prod, &quo, &m);
    c->g = prod;
  }
  c->nroots = 2 * t;
  for (unsigned power = 1; power < 2 * t + 1; power++) c->roots[power - 1] = c->gf.exp[power % c->gf.size];
  return finish_code(c);
}

/* rs.h:18-39 */
int oracle_rs_init(code_t *c, unsigned q, unsigned t, unsigned mu, unsigned step) {
  memset(c, 0, sizeof(*c));
  if (oracle_gf_init(&c->gf, q, 0)) return -1;
  c->family = 1;
  c->n = c->gf.n;
  c->t = t;
  pset1(&c->g, 1);
  for (unsigned i = 0; i < 2 * t; i++) {
    poly_t lin, prod;
    unsigned root = c->gf.exp[(mu + i * step) % c->gf.size];
    lin.len = 2;
    lin.c[0] = (uint16_t)root;
    lin.c[1] = 1;
    pmul(&c->gf, &prod, &c->g, &lin);
    c->g = prod;
    c->roots[i] = (uint16_t)root;
  }
  c->nroots = 2 * t;
  return finish_code(c);
}

code_t *oracle_code_new(int family, unsigned q, unsigned t, unsigned mu, unsigned step) {
  code_t *c = (code_t *)malloc(sizeof(code_t));
  int rc = family == 0 ? oracle_bch_init(c, q, t) : oracle_rs_init(c, q, t, mu, step);
  if (rc) { free(c); return 0; }
  return c;
}
void oracle_code_free(code_t *c) { free(c); }

/* out7 = n, l, k, dmin, t, deg g, deg h */
void oracle_code_params(const code_t *c, unsigned *out7, double *rate) {
  out7[0] = c->n; out7[1] = c->l; out7[2] = c->k; out7[3] = c->dmin; out7[4] = c->t;
  out7[5] = (unsigned)pdeg(&c->g); out7[6] = (unsigned)pdeg(&c->h);
  *rate = c->rate;
}
int oracle_code_poly(const code_t *c, int which, uint32_t *out) {
  const poly_t *p = which ? &c->h : &c->g;
  for (int i = 0; i < p->len; i++) out[i] = p->c[i];
  return p->len;
}

/* cyclic.h:346-359: row0 = h reversed (h_l .. h_0) then zeros; row r = row0 rotated right r.
 * k rows for H(); `rows` lets the caller ask for the redundant n-row variant (extension). */
void oracle_code_H(const code_t *c, unsigned rows, uint8_t *out) {
  const unsigned n = c->n;
  memset(out, 0, (size_t)rows * n);
  for (int i = 0; i < c->h.len; i++) out[i] = (uint8_t)c->h.c[c->h.len - 1 - i];
  for (unsigned r = 1; r < rows; r++)
    for (unsigned col = 0; col < n; col++) out[r * n + (col + 1) % n] = out[(r - 1) * n + col];
}

/* cyclic.h:35-40, :289-311 systematic encode: a*x^k + (a*x^k mod g) */
void oracle_encode(const code_t *c, const uint8_t *msg, uint8_t *word) {
  poly_t xk, rem;
  xk.len = (int)c->n;
  memset(xk.c, 0, sizeof(uint16_t) * xk.len);
  for (unsigned i = 0; i < c->l; i++) xk.c[c->k + i] = msg[i];
  pdivmod(&c->gf, 0, &rem, &xk, &c->g);
  for (unsigned i = 0; i < c->n; i++) word[i] = (uint8_t)(xk.c[i] ^ (i < (unsigned)rem.len && i < c->k ? rem.c[i] : 0));
}

/* cyclic.h:53-63 */
static int syndromes_of(const code_t *c, const poly_t *b, uint16_t *s) {
  int any = 0;
  for (unsigned j = 0; j < c->nroots; j++) {
    s[j] = (uint16_t)peval(&c->gf, b, c->roots[j]);
    any |= s[j] != 0;
  }
  return any;
}
void oracle_syndromes(const code_t *c, const uint8_t *word, uint16_t *s) {
  poly_t b;
  b.len = (int)c->n;
  for (unsigned i = 0; i < c->n; i++) b.c[i] = word[i];
  syndromes_of(c, &b, s);
}

/* naive error values, rs.h:41-78: solve  sum_m y_m X_m^(i+1) = s_i , i = 0..v-1.
 * (Gauss-Jordan; the reference's polynomial-row solver of linear_equation_system.h:12-88 returns
 * the same unique solution whenever the X_m are distinct and non-zero.)  returns -1 if singular */
static int solve_values(const gf_t *gf, unsigned v, const uint16_t *X, const uint16_t *s, uint16_t *y) {
  uint16_t A[64][65];
  if (v > 64) return -1;
  for (unsigned i = 0; i < v; i++) {
    for (unsigned m = 0; m < v; m++) {
      unsigned p = 1;
      for (unsigned e = 0; e <= i; e++) p = gmul(gf, p, X[m]);
      A[i][m] = (uint16_t)p;
    }
    A[i][v] = s[i];
  }
  for (unsigned col = 0; col < v; col++) {
    unsigned piv = col;
    while (piv < v && !A[piv][col]) piv++;
    if (piv == v) return -1;
    if (piv != col)
      for (unsigned j = 0; j <= v; j++) { uint16_t tmp = A[piv][j]; A[piv][j] = A[col][j]; A[col][j] = tmp; }
    unsigned inv = ginv(gf, A[col][col]);
    for (unsigned j = 0; j <= v; j++) A[col][j] = (uint16_t)gmul(gf, A[col][j], inv);
    for (unsigned i = 0; i < v; i++)
      if (i != col && A[i][col]) {
        unsigned f = A[i][col];
        for (unsigned j = 0; j <= v; j++) A[i][j] ^= (uint16_t)gmul(gf, A[col][j], f);
      }
  }
  for (unsigned m = 0; m < v; m++) y[m] = A[m][v];
  return 0;
}

/* cyclic.h:207-252 with Algorithm = euklid_tag.
 * returns 0 ok (out = corrected word, *nerr = number of corrected positions),
 *         1 decoding_failure, 2 any other exception of the reference. */
int oracle_hard_correct(const code_t *c, const uint8_t *word, const uint32_t *erasures,
                        unsigned n_erasures, uint8_t *out, unsigned *nerr) {
  const gf_t *gf = &c->gf;
  poly_t b;
  uint16_t s[512];
  b.len = (int)c->n;
  for (unsigned i = 0; i < c->n; i++) {
    if (word[i] & ~c->gf.n) return 2; /* galois.h:162-165 "Value is not an element of the field" */
    b.c[i] = word[i];
  }
  *nerr = 0;
  if (syndromes_of(c, &b, s)) {
    /* hard_decision.h:157-196 */
    const unsigned fk = c->nroots / 2;
    const int max = (int)((2 * fk + n_erasures) / 2);
    poly_t u, S, lin, tmp, r_prev, r_cur, w_prev, w_cur, quo, rem, prod, lambda;
    pset1(&u, 1);
    for (unsigned e = 0; e < n_erasures; e++) {
      lin.len = 2;
      lin.c[0] = 1;
      lin.c[1] = (uint16_t)gf->exp[erasures[e] % gf->size];
      pmul(gf, &tmp, &u, &lin);
      u = tmp;
    }
    S.len = (int)c->nroots;
    for (unsigned j = 0; j < c->nroots; j++) S.c[j] = s[j];
    pmul(gf, &r_prev, &S, &u);                 /* r[0] = s * u */
    r_cur.len = (int)(2 * fk + 1);             /* r[1] = x^(dmin-1) */
    memset(r_cur.c, 0, sizeof(uint16_t) * r_cur.len);
    r_cur.c[2 * fk] = 1;
    w_prev = u;                                /* w[0] = u, w[1] = 0 */
    pset1(&w_cur, 0);
    while (pdeg(&r_cur) >= max) {
      if (pdivmod(gf, &quo, &rem, &r_prev, &r_cur)) return 2;
      pmul(gf, &prod, &quo, &w_cur);
      tmp = w_prev;
      padd(&tmp, &prod);                       /* w[i-1] + q * w[i] */
      w_prev = w_cur;
      w_cur = tmp;
      r_prev = r_cur;
      r_cur = rem;
    }
    if (w_cur.len == 0 || w_cur.c[0] == 0) return 1; /* "Cannot invert last element" */
    {
      unsigned inv = ginv(gf, w_cur.c[0]);
      lambda = w_cur;
      for (int i = 0; i < lambda.len; i++) lambda.c[i] = (uint16_t)gmul(gf, lambda.c[i], inv);
    }
    /* lambda.reverse(): reverse coefficients [0, degree] (polynomial.h:176-179) */
    int dl = pdeg(&lambda);
    for (int i = 0, j = dl; i < j; i++, j--) { uint16_t x = lambda.c[i]; lambda.c[i] = lambda.c[j]; lambda.c[j] = x; }
    /* zeroes(): cyclic.h:126-150 -- brute-force roots over the non-zero elements, sorted by power */
    uint16_t zero[512];
    unsigned nz = 0;
    for (unsigned p = 0; p < gf->n; p++) {
      unsigned el = gf->exp[p];
      if (peval(gf, &lambda, el) == 0) zero[nz++] = (uint16_t)el;
    }
    if ((int)nz != dl) return 1; /* "Sigma(x) has to have .. zeroes" */
    if (nz == 0) return 1;       /* none_of over an empty range -> "0 is zero in Sigma(x)" */
    /* error values */
    uint16_t val[512];
    if (c->family == 0) {
      for (unsigned m = 0; m < nz; m++) val[m] = 1; /* bch.h:80-83 */
    } else {
      if (nz > c->nroots) return 2;
      if (solve_values(gf, nz, zero, s, val)) return 2; /* "Linear equation system not solvable" */
    }
    for (unsigned m = 0; m < nz; m++) {
      unsigned pos = gf->log[zero[m]]; /* error_positions: cyclic.h:152-159 */
      b.c[pos] ^= val[m];
    }
    *nerr = nz;
    if (syndromes_of(c, &b, s)) return 1; /* "Corrected word is not a codeword" */
  }
  for (unsigned i = 0; i < c->n; i++) out[i] = (uint8_t)b.c[i];
  return 0;
}

int oracle_hard_correct_batch(const code_t *c, const uint8_t *words, uint64_t count,
                              const uint32_t *erasures, unsigned n_erasures, uint8_t *out,
                              uint32_t *nerr, uint8_t *status) {
  for (uint64_t w = 0; w < count; w++) {
    unsigned ne = 0;
    memset(out + w * c->n, 0, c->n);
    status[w] = (uint8_t)oracle_hard_correct(c, words + w * c->n, erasures, n_erasures, out + w * c->n, &ne);
    nerr[w] = ne;
  }
  return 0;
}
