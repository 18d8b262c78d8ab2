// amwg_ld.cuh -- the ld.* log densities as device functions, operation order as in
// /root/reference/distributions.js (line numbers cited per function).  fp64, no FMA contraction.
#pragma once
#include "amwg_math.cuh"

namespace amwg {

#define AMWG_JS_PI 3.141592653589793

// distributions.js:63-76  -- 6-term Lanczos (NOT libm lgamma)
__device__ __forceinline__ double ld_lgamma(double x) {
  const double cof[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                         -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5};
  double ser = 1.000000000190015, xx = x, y = x, tmp = x + 5.5;
  tmp -= (xx + 0.5) * js_log(tmp);
#pragma unroll
  for (int j = 0; j < 6; j++) { y += 1.0; ser += cof[j] / y; }
  return js_log(2.5066282746310005 * ser / xx) - tmp;
}
// :79-81
__device__ __forceinline__ double ld_lfactorial(double n) { return n < 0 ? CUDART_NAN : ld_lgamma(n + 1); }
// :84-86
__device__ __forceinline__ double ld_lchoose(double n, double k) { return ld_lfactorial(n) - ld_lfactorial(k) - ld_lfactorial(n - k); }
// :89-91
__device__ __forceinline__ double ld_lbeta(double a, double b) { return ld_lgamma(a) + ld_lgamma(b) - ld_lgamma(a + b); }

// :104-113
__device__ __forceinline__ double ld_beta(double x, double shape1, double shape2) {
  if (x > 1 || x < 0) return -CUDART_INF;
  if (shape1 == 1 && shape2 == 1) return 0;
  return (shape1 - 1) * js_log(x) + (shape2 - 1) * js_log(1 - x) - ld_lbeta(shape1, shape2);
}
// :115-117
__device__ __forceinline__ double ld_cauchy(double x, double location, double scale) {
  return js_log(scale) - js_log(js_pow(x - location, 2) + js_pow(scale, 2)) - js_log(AMWG_JS_PI);
}
// :119-121
__device__ __forceinline__ double ld_norm(double x, double mean, double sd) {
  return -0.5 * js_log(2 * AMWG_JS_PI) - js_log(sd) - js_pow(x - mean, 2) / (2 * sd * sd);
}
// :136-138
__device__ __forceinline__ double ld_laplace(double x, double location, double scale) {
  return (-fabs(x - location) / scale) - js_log(2 * scale);
}
// :142-152
__device__ __forceinline__ double ld_gamma(double x, double shape, double rate) {
  double scale = 1 / rate;
  if (x < 0) return -CUDART_INF;
  if (x == 0 && shape == 1) return -js_log(scale);
  return (shape - 1) * js_log(x) - x / scale - ld_lgamma(shape) - shape * js_log(scale);
}
// :154-159
__device__ __forceinline__ double ld_invgamma(double x, double shape, double scale) {
  if (x <= 0) return -CUDART_INF;
  return -(shape + 1) * js_log(x) - scale / x - ld_lgamma(shape) + shape * js_log(scale);
}
// :161-167
__device__ __forceinline__ double ld_lnorm(double x, double meanlog, double sdlog) {
  if (x <= 0) return -CUDART_INF;
  return -js_log(x) - 0.5 * js_log(2 * AMWG_JS_PI) - js_log(sdlog) - js_pow(js_log(x) - meanlog, 2) / (2 * sdlog * sdlog);
}
// :169-174
__device__ __forceinline__ double ld_pareto(double x, double scale, double shape) {
  if (x < scale) return -CUDART_INF;
  return js_log(shape) + shape * js_log(scale) - (shape + 1) * js_log(x);
}
// :176-180
__device__ __forceinline__ double ld_t(double x, double location, double scale, double df) {
  df = df > 1e100 ? 1e100 : df;
  return ld_lgamma((df + 1) / 2) - ld_lgamma(df / 2) - js_log(sqrt(AMWG_JS_PI * df) * scale) +
         js_log(js_pow(1 + (1 / df) * js_pow((x - location) / scale, 2), -(df + 1) / 2));
}
// :185-191
__device__ __forceinline__ double ld_weibull(double x, double shape, double scale) {
  if (x < 0) return -CUDART_INF;
  if (x == 0 && shape < 1) return CUDART_INF;
  double tmp1 = js_pow(x / scale, shape - 1);
  double tmp2 = tmp1 * (x / scale);
  return -tmp2 + js_log(shape * tmp1 / scale);
}
// :196-201
__device__ __forceinline__ double ld_logis(double x, double location, double scale) {
  x = fabs((x - location) / scale);
  double e = js_exp(-x), f = 1.0 + e;
  return -(x + js_log(scale * f * f));
}
// :217-219
__device__ __forceinline__ double ld_exp(double x, double rate) { return x < 0 ? -CUDART_INF : js_log(rate) - rate * x; }
// :221-223
__device__ __forceinline__ double ld_unif(double x, double mn, double mx) {
  return (x < mn || x > mx) ? -CUDART_INF : js_log(1 / (mx - mn));
}
// :228-230
__device__ __forceinline__ double ld_bern(double x, double prob) {
  return !(x == 0 || x == 1) ? -CUDART_INF : js_log(x * prob + (1 - x) * (1 - prob));
}
// :240-248
__device__ __forceinline__ double ld_binom(double x, double size, double prob) {
  if (x > size || x < 0) return -CUDART_INF;
  if (prob == 0 || prob == 1) return (size * prob) == x ? 0 : -CUDART_INF;
  return ld_lchoose(size, x) + x * js_log(prob) + (size - x) * js_log(1 - prob);
}
// :267-272
__device__ __forceinline__ double ld_nbinom(double x, double size, double prob) {
  if (x < 0) return -CUDART_INF;
  return ld_lchoose(x + size - 1, size - 1) + x * js_log(1 - prob) + size * js_log(prob);
}
// :274-280
__device__ __forceinline__ double ld_hyper(double x, double m, double n, double k) {
  if (x < 0 || x > k) return -CUDART_INF;
  return ld_lchoose(m, x) + ld_lchoose(n, k - x) - ld_lchoose(m + n, k);
}
// :282-284
__device__ __forceinline__ double ld_pois(double x, double lambda) {
  return x < 0 ? -CUDART_INF : js_log(lambda) * x - lambda - ld_lfactorial(x);
}

}  // namespace amwg
