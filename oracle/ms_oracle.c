/* oracle/ms_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's iterative soft-decision decoder
 * (hannesweisbach/channelcoding, src/codes/soft_decision.h) on a dense 0/1 parity-check
 * matrix.  It keeps the reference's dense k x n loops, evaluation order and float32/double
 * types so that every output is bit-identical to REF-FIXED (= reference + the one-token
 * matrix.h:50 end() fix); tests/test_oracle_pin.py checks exactly that against
 * oracle/_ref/libccref.so and the committed tests/golden/ fixtures.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this file.
 * Build with -ffp-contract=off (no FMA contraction).
 *
 * Parity status:  MS, NMS, OMS, SCMS1, SCMS2, 2DNMS with stop rule REF_ZERO_OVERLAP: pinned.
 *                 stop rules GF2_PARITY / NONE, runtime alpha/beta/max_iter outside the tag
 *                 instantiations of oracle/ref_shim.cc, and SPA (sum-product): PARITY UNPINNED
 *                 (no reference implementation exists; they reuse the pinned loop structure).
 *                 Fixed-point min-sum (oracle_min_sum_fixed, variants MS_Q / NMS_Q / OMS_Q): PARITY UNPINNED --
 *                 the reference has no integer decoder; the function below is min_sum__ with Q = R = int and
 *                 is cross-checked against the pinned float decoder for a fine quantiser
 *                 (tests/test_oracle_golden.py::test_fixed_point_converges_to_float).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { V_MS = 0, V_NMS = 1, V_OMS = 2, V_SCMS1 = 3, V_SCMS2 = 4, V_NMS2D = 5, V_SPA = 6 };
enum { STOP_REF_ZERO_OVERLAP = 0, STOP_GF2_PARITY = 1, STOP_NONE = 2 };

/* soft_decision.h:75-77 */
static int signum_f(float v) { return (0.0f < v) - (v < 0.0f); }

/* soft_decision.h:86-98  column_sum: rows ascending, starting from R(0) */
static void column_sum(const uint8_t *H, unsigned rows, unsigned cols, const float *r, float *out) {
  for (unsigned c = 0; c < cols; c++) out[c] = 0.0f;
  for (unsigned row = 0; row < rows; row++)
    for (unsigned col = 0; col < cols; col++)
      if (H[row * cols + col]) out[col] += r[row * cols + col];
}

/* check-node functor fn_h(min) followed by static_cast<R>(sign * fn(min))
 * (soft_decision.h:118 with the functors of :204-213, :245-251) */
static float cn_value(int variant, int sign, float min, double alpha, double beta) {
  switch (variant) {
  case V_NMS:
  case V_NMS2D: {
    /* normalised_horizontal<float>(min, alpha): alpha is bound as double and converted to
     * const float& at the call (:211-213, :233-236, :289-294) */
    float a = (float)alpha;
    float v = a * min;
    return (float)((float)sign * v);
  }
  case V_OMS: {
    /* lambda returns std::max(min - beta, Result_t(0)) with Result_t = double (:245-251);
     * sign * double is double, static_cast<float> at :118 */
    double d = (double)min - beta;
    if (!(d > 0.0)) d = 0.0; /* std::max(d, 0.0): returns d only if 0.0 < d */
    return (float)((double)sign * d);
  }
  default: /* unmodified_horizontal (:204) */
    return (float)((float)sign * min);
  }
}

/* symbol-node functor fn_v(exclusive_colsum, y, q_old)  (soft_decision.h:205-218, :261-281) */
static float vn_value(int variant, float e, float y, float q_old, double beta) {
  switch (variant) {
  case V_SCMS1: {
    float tmp = e + y;
    if (signum_f(q_old) == 0 || signum_f(q_old) == signum_f(tmp)) return tmp;
    return 0.0f;
  }
  case V_SCMS2: {
    float tmp = e + y;
    if (tmp * q_old > 0) return tmp;
    return 0.5f * (tmp + q_old);
  }
  case V_NMS2D: {
    float b = (float)beta;
    float p = b * e; /* separate multiply and add: no contraction */
    return p + y;
  }
  default: /* unmodified_vertical (:205-209) */
    return e + y;
  }
}

/* Returns 0 on success (stop test passed at 0-based iteration *iter_out), 1 on
 * decoding_failure (soft_decision.h:201; then *iter_out = max_iter and b/L hold the state of
 * the last iteration -- the reference itself returns nothing in that case).
 * H: rows x cols row-major 0/1.  y: cols channel values.  llr_scale only used by V_SPA. */
int oracle_min_sum(const uint8_t *H, unsigned rows, unsigned cols, const float *y, int variant,
                   double alpha, double beta, unsigned max_iter, int stop_rule, uint8_t *b_out,
                   float *L_out, unsigned *iter_out) {
  const size_t E = (size_t)rows * cols;
  float *q = (float *)calloc(E, sizeof(float));   /* matrix<Q> q(rows, cols) value-init (:167) */
  float *r = (float *)calloc(E, sizeof(float));   /* matrix<R> r (:168) */
  float *col_sums = (float *)calloc(cols, sizeof(float));
  float *L = (float *)calloc(cols, sizeof(float));
  uint8_t *b = (uint8_t *)calloc(cols, 1);
  int rc = 1;
  unsigned iteration;
  for (iteration = 0; iteration < max_iter; iteration++) {
    /* vertical__ (:125-140) */
    column_sum(H, rows, cols, r, col_sums);
    for (unsigned row = 0; row < rows; row++)
      for (unsigned col = 0; col < cols; col++)
        if (H[row * cols + col]) {
          const float e = col_sums[col] - r[row * cols + col];
          q[row * cols + col] = vn_value(variant, e, y[col], q[row * cols + col], beta);
        }
    /* horizontal__ (:101-122) */
    for (unsigned row = 0; row < rows; row++)
      for (unsigned col = 0; col < cols; col++)
        if (H[row * cols + col]) {
          if (variant == V_SPA) {
            /* extension (unpinned): r = 2 atanh( prod_{i != col} tanh(q_i / 2) ) in float.  The exclusive product is
             * associated like the device kernel's (ms_cyclic.cuh: prefix of the edges before col, in ascending column
             * order, times the suffix accumulated from the last edge downwards), so that the only difference left
             * between this restatement and the kernel is the last-ulp behaviour of tanhf / atanhf (libm vs CUDA). */
            float prefix = 1.0f, suffix = 1.0f;
            for (unsigned i = 0; i < col; i++)
              if (H[row * cols + i]) prefix *= tanhf(0.5f * q[row * cols + i]);
            for (unsigned i = cols; i-- > col + 1;)
              if (H[row * cols + i]) suffix *= tanhf(0.5f * q[row * cols + i]);
            float prod = prefix * suffix;
            /* clamp like the device kernel so that atanh stays finite */
            const float lim = 0.99999994f;
            if (prod > lim) prod = lim;
            if (prod < -lim) prod = -lim;
            r[row * cols + col] = 2.0f * atanhf(prod);
            continue;
          }
          int sign = 1;
          float min = FLT_MAX;
          for (unsigned i = 0; i < cols; i++)
            if (i != col && H[row * cols + i]) {
              sign *= signum_f(q[row * cols + i]);
              const float a = fabsf(q[row * cols + i]);
              if (a < min) min = a; /* std::min(min, a) */
            }
          r[row * cols + col] = cn_value(variant, sign, min, alpha, beta);
        }
    /* totals and hard decision (:178-183, codes.h:43-52) */
    column_sum(H, rows, cols, r, col_sums);
    for (unsigned c = 0; c < cols; c++) {
      L[c] = col_sums[c] + y[c];
      b[c] = (uint8_t)(L[c] < 0);
    }
    /* stop test: syndrome(H, b) (:79-84) with matrix::operator* (matrix.h:57-67) accumulating
     * in uint8_t; a check passes iff the integer overlap is 0 mod 256 */
    int stop = 0;
    if (stop_rule == STOP_REF_ZERO_OVERLAP) {
      stop = 1;
      for (unsigned row = 0; row < rows && stop; row++) {
        uint8_t acc = 0;
        for (unsigned c = 0; c < cols; c++) acc = (uint8_t)(acc + H[row * cols + c] * b[c]);
        if (acc) stop = 0;
      }
    } else if (stop_rule == STOP_GF2_PARITY) {
      stop = 1;
      for (unsigned row = 0; row < rows && stop; row++) {
        unsigned acc = 0;
        for (unsigned c = 0; c < cols; c++) acc ^= (unsigned)(H[row * cols + c] & b[c]);
        if (acc) stop = 0;
      }
    } else {
      stop = (iteration + 1 == max_iter); /* STOP_NONE: run all iterations, never fail */
    }
    if (stop) {
      rc = 0;
      break;
    }
  }
  memcpy(b_out, b, cols);
  if (L_out) memcpy(L_out, L, cols * sizeof(float));
  *iter_out = iteration; /* == max_iter on failure */
  free(q); free(r); free(col_sums); free(L); free(b);
  return rc;
}

/* batch helper: frames x cols inputs; failed[f] = return code */
int oracle_min_sum_batch(const uint8_t *H, unsigned rows, unsigned cols, const float *y,
                         uint64_t frames, int variant, double alpha, double beta,
                         unsigned max_iter, int stop_rule, uint8_t *bits, float *L,
                         uint32_t *iter, uint8_t *failed) {
  for (uint64_t f = 0; f < frames; f++) {
    unsigned it = 0;
    int rc = oracle_min_sum(H, rows, cols, y + f * cols, variant, alpha, beta, max_iter, stop_rule,
                            bits + f * cols, L ? L + f * cols : 0, &it);
    iter[f] = it;
    failed[f] = (uint8_t)rc;
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Fixed-point min-sum (north_star "fixed-point min-sum"; include/ccgpu.h CCGPU_MS_Q / NMS_Q / OMS_Q).
 * PARITY UNPINNED: the reference has no integer instantiation.  This is min_sum__
 * (soft_decision.h:161-202) with its loops, evaluation structure and three-valued signum kept and
 * Q = R = int:
 *   quantiser      y_i = clamp(rint(y * q_scale), -q_y_max, +q_y_max)   float32 product, ties to even,
 *                  NaN -> 0 (the channel values enter the loop exactly once, here)
 *   vertical__     q = (S_c - r) + y_c in wide accumulators (:135-136), no saturation
 *   horizontal__   sign = prod signum(q_i), min = min |q_i| over the others (:106-116);
 *                  r = sign * fn_h(min(min, q_msg_max))  -- the message saturates before fn_h
 *                    MS_Q   fn_h(m) = m                                       (:204)
 *                    NMS_Q  fn_h(m) = rne(A * m / 1024), A = rint(alpha*1024) (:211-213 in Q10)
 *                    OMS_Q  fn_h(m) = max(m - B, 0),     B = rint(beta * q_scale)  (:245-251)
 *   totals         L_c = S_c + y_c, b_c = L_c < 0 (:178-183), stop rules as above (:79-84)
 * Integer addition is associative, so the summation order of column_sum is irrelevant here. */
enum { V_MS_Q = 7, V_NMS_Q = 8, V_OMS_Q = 9 };

static int signum_i(int v) { return (0 < v) - (v < 0); }

int oracle_quantise(float y, float q_scale, int q_y_max) {
  const float t = y * q_scale;
  if (t != t) return 0;
  if (t >= (float)q_y_max) return q_y_max;
  if (t <= -(float)q_y_max) return -q_y_max;
  return (int)lrintf(t); /* default rounding mode: to nearest, ties to even */
}

static long long rne_shift10(long long t) { /* t >= 0: round(t / 1024) to nearest, ties to even */
  const long long fl = t >> 10, rem = t & 1023;
  if (rem > 512 || (rem == 512 && (fl & 1))) return fl + 1;
  return fl;
}

static int cn_value_fixed(int variant, int sign, int min, int q_msg_max, int A, int B) {
  int m = min < q_msg_max ? min : q_msg_max;
  if (variant == V_NMS_Q) m = (int)rne_shift10((long long)A * m);
  else if (variant == V_OMS_Q) m = m - B > 0 ? m - B : 0;
  return sign * m;
}

int oracle_min_sum_fixed(const uint8_t *H, unsigned rows, unsigned cols, const float *y, int variant,
                         double alpha, double beta, unsigned max_iter, int stop_rule, float q_scale,
                         int q_y_max, int q_msg_max, uint8_t *b_out, int32_t *L_out, unsigned *iter_out) {
  const size_t E = (size_t)rows * cols;
  int *q = (int *)calloc(E, sizeof(int));
  int *r = (int *)calloc(E, sizeof(int));
  int *yi = (int *)calloc(cols, sizeof(int));
  int *col_sums = (int *)calloc(cols, sizeof(int));
  int *L = (int *)calloc(cols, sizeof(int));
  uint8_t *b = (uint8_t *)calloc(cols, 1);
  const int A = (int)lrint(alpha * 1024.0), B = (int)lrint(beta * (double)q_scale);
  int rc = 1;
  unsigned iteration;
  for (unsigned c = 0; c < cols; c++) yi[c] = oracle_quantise(y[c], q_scale, q_y_max);
  for (iteration = 0; iteration < max_iter; iteration++) {
    /* vertical__ (:125-140) */
    for (unsigned c = 0; c < cols; c++) col_sums[c] = 0;
    for (unsigned row = 0; row < rows; row++)
      for (unsigned col = 0; col < cols; col++)
        if (H[row * cols + col]) col_sums[col] += r[row * cols + col];
    for (unsigned row = 0; row < rows; row++)
      for (unsigned col = 0; col < cols; col++)
        if (H[row * cols + col]) q[row * cols + col] = (col_sums[col] - r[row * cols + col]) + yi[col];
    /* horizontal__ (:101-122) */
    for (unsigned row = 0; row < rows; row++)
      for (unsigned col = 0; col < cols; col++)
        if (H[row * cols + col]) {
          int sign = 1, min = INT32_MAX;
          for (unsigned i = 0; i < cols; i++)
            if (i != col && H[row * cols + i]) {
              sign *= signum_i(q[row * cols + i]);
              const int a = abs(q[row * cols + i]);
              if (a < min) min = a;
            }
          r[row * cols + col] = cn_value_fixed(variant, sign, min, q_msg_max, A, B);
        }
    /* totals and hard decision (:178-183) */
    for (unsigned c = 0; c < cols; c++) col_sums[c] = 0;
    for (unsigned row = 0; row < rows; row++)
      for (unsigned col = 0; col < cols; col++)
        if (H[row * cols + col]) col_sums[col] += r[row * cols + col];
    for (unsigned c = 0; c < cols; c++) {
      L[c] = col_sums[c] + yi[c];
      b[c] = (uint8_t)(L[c] < 0);
    }
    int stop = 0;
    if (stop_rule == STOP_REF_ZERO_OVERLAP) {
      stop = 1;
      for (unsigned row = 0; row < rows && stop; row++) {
        uint8_t acc = 0;
        for (unsigned c = 0; c < cols; c++) acc = (uint8_t)(acc + H[row * cols + c] * b[c]);
        if (acc) stop = 0;
      }
    } else if (stop_rule == STOP_GF2_PARITY) {
      stop = 1;
      for (unsigned row = 0; row < rows && stop; row++) {
        unsigned acc = 0;
        for (unsigned c = 0; c < cols; c++) acc ^= (unsigned)(H[row * cols + c] & b[c]);
        if (acc) stop = 0;
      }
    } else {
      stop = (iteration + 1 == max_iter);
    }
    if (stop) {
      rc = 0;
      break;
    }
  }
  memcpy(b_out, b, cols);
  if (L_out) memcpy(L_out, L, cols * sizeof(int32_t));
  *iter_out = iteration;
  free(q); free(r); free(yi); free(col_sums); free(L); free(b);
  return rc;
}

int oracle_min_sum_fixed_batch(const uint8_t *H, unsigned rows, unsigned cols, const float *y, uint64_t frames,
                               int variant, double alpha, double beta, unsigned max_iter, int stop_rule,
                               float q_scale, int q_y_max, int q_msg_max, uint8_t *bits, int32_t *L,
                               uint32_t *iter, uint8_t *failed) {
  for (uint64_t f = 0; f < frames; f++) {
    unsigned it = 0;
    int rc = oracle_min_sum_fixed(H, rows, cols, y + f * cols, variant, alpha, beta, max_iter, stop_rule, q_scale,
                                  q_y_max, q_msg_max, bits + f * cols, L ? L + f * cols : 0, &it);
    iter[f] = it;
    failed[f] = (uint8_t)rc;
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Sum-product in DOUBLE precision: the yardstick for the float32 sum-product kernel (extension, parity unpinned --
 * the reference has no tanh-rule decoder).  Same flooding loop (soft_decision.h:161-202), same clamp of the exclusive
 * product (the float32 constant 0.99999994), every operation in double.  L_out receives the totals of the last
 * executed iteration as doubles. */
int oracle_spa_f64(const uint8_t *H, unsigned rows, unsigned cols, const float *y, unsigned max_iter, int stop_rule,
                   uint8_t *b_out, double *L_out, unsigned *iter_out) {
  const size_t E = (size_t)rows * cols;
  double *q = (double *)calloc(E, sizeof(double));
  double *r = (double *)calloc(E, sizeof(double));
  double *cs = (double *)calloc(cols, sizeof(double));
  double *L = (double *)calloc(cols, sizeof(double));
  uint8_t *b = (uint8_t *)calloc(cols, 1);
  const double lim = (double)0.99999994f;
  int rc = 1;
  unsigned iteration;
  for (iteration = 0; iteration < max_iter; iteration++) {
    for (unsigned c = 0; c < cols; c++) cs[c] = 0.0;
    for (unsigned row = 0; row < rows; row++)
      for (unsigned col = 0; col < cols; col++)
        if (H[row * cols + col]) cs[col] += r[row * cols + col];
    for (unsigned row = 0; row < rows; row++)
      for (unsigned col = 0; col < cols; col++)
        if (H[row * cols + col]) q[row * cols + col] = (cs[col] - r[row * cols + col]) + (double)y[col];
    for (unsigned row = 0; row < rows; row++)
      for (unsigned col = 0; col < cols; col++)
        if (H[row * cols + col]) {
          double prod = 1.0;
          for (unsigned i = 0; i < cols; i++)
            if (i != col && H[row * cols + i]) prod *= tanh(0.5 * q[row * cols + i]);
          if (prod > lim) prod = lim;
          if (prod < -lim) prod = -lim;
          r[row * cols + col] = 2.0 * atanh(prod);
        }
    for (unsigned c = 0; c < cols; c++) cs[c] = 0.0;
    for (unsigned row = 0; row < rows; row++)
      for (unsigned col = 0; col < cols; col++)
        if (H[row * cols + col]) cs[col] += r[row * cols + col];
    for (unsigned c = 0; c < cols; c++) {
      L[c] = cs[c] + (double)y[c];
      b[c] = (uint8_t)(L[c] < 0);
    }
    int stop = 0;
    if (stop_rule == STOP_REF_ZERO_OVERLAP) {
      stop = 1;
      for (unsigned row = 0; row < rows && stop; row++) {
        uint8_t acc = 0;
        for (unsigned c = 0; c < cols; c++) acc = (uint8_t)(acc + H[row * cols + c] * b[c]);
        if (acc) stop = 0;
      }
    } else if (stop_rule == STOP_GF2_PARITY) {
      stop = 1;
      for (unsigned row = 0; row < rows && stop; row++) {
        unsigned acc = 0;
        for (unsigned c = 0; c < cols; c++) acc ^= (unsigned)(H[row * cols + c] & b[c]);
        if (acc) stop = 0;
      }
    } else {
      stop = (iteration + 1 == max_iter);
    }
    if (stop) {
      rc = 0;
      break;
    }
  }
  memcpy(b_out, b, cols);
  if (L_out) memcpy(L_out, L, cols * sizeof(double));
  *iter_out = iteration;
  free(q); free(r); free(cs); free(L); free(b);
  return rc;
}

int oracle_spa_f64_batch(const uint8_t *H, unsigned rows, unsigned cols, const float *y, uint64_t frames,
                         unsigned max_iter, int stop_rule, uint8_t *bits, double *L, uint32_t *iter, uint8_t *failed) {
  for (uint64_t f = 0; f < frames; f++) {
    unsigned it = 0;
    int rc = oracle_spa_f64(H, rows, cols, y + f * cols, max_iter, stop_rule, bits + f * cols, L ? L + f * cols : 0, &it);
    iter[f] = it;
    failed[f] = (uint8_t)rc;
  }
  return 0;
}
