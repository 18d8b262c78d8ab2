// amwg_jit_full_kernel.cuh -- the run-time specialised form of amwg_sweep_kernel<false>: models whose every step evaluates the whole
// log_post (no term cache, no pre-evaluated statistics) -- models with binary parameters, bit-faithful (`faithful`) lowerings, small
// models in general. Same reference path (mcmc.js:985-1039, 886-892, 685-688, 517-553, 753-767, 43-54 + ld.*), same operations in
// the same order as the interpreter kernel -- the generated jit_logpost() is the program's instructions printed as straight-line
// CUDA C++ (csrc/amwg_jit.cuh), so the draws are bit-identical to the interpreter's and to the reference's -- minus the interpreter:
// no dispatch, no operand decoding, plates and loops inlined with their sizes and columns as constants.
//
// The generated part of the translation unit comes first (J* constants, parameter tables, jit_logpost / jit_derived).
#pragma once

namespace amwg {

struct JitArgs {
  ChainArrays a;
  SweepArgs sa;
  const double* col[JMAXCOL];        // the model's data columns in HBM
  const unsigned char* adapting;     // [JD] host-maintained (start/stop_adaptation)
};

#define ST(c) sp[(unsigned long long)(c) * ss]
// state component c as an evaluation sees it: the proposal for the moved one
#define CM(c) (((c) == moved) ? val : ST(c))

// sum_i ld.bern(y_i, p): sequential, bit-faithful to distributions.js:228-230 (the interpreter's plate_bern_iid)
__device__ __forceinline__ double jit_plate_bern(unsigned saddr, int n, double p, double lp) {
  const double l1 = js_log(1.0 * p + (1 - 1.0) * (1 - p));
  const double l0 = js_log(0.0 * p + (1 - 0.0) * (1 - p));
#pragma unroll 4
  for (int i = 0; i < n; ++i) {
    const double yi = lds_f64_sa(saddr + 8u * (unsigned)i);
    lp = lp + (yi == 1.0 ? l1 : (yi == 0.0 ? l0 : -CUDART_INF));
  }
  return lp;
}
// The same sum with the column read as a bit mask (built once per launch, below): per point one bit test, a select and the add --
// the adds stay sequential and in data order (each one rounds), which is all that bit-faithfulness asks for.
template <int N>
__device__ __forceinline__ double jit_plate_bern_mask(const unsigned char* smem, unsigned data_off, unsigned mask_off, double p, double lp) {
  const unsigned* __restrict__ mask = reinterpret_cast<const unsigned*>(smem + mask_off);
  if (mask[(N + 31) / 32] != 0u) return jit_plate_bern(smem_u32(smem) + data_off, N, p, lp);   // a point that is neither 0 nor 1
  const double l1 = js_log(1.0 * p + (1 - 1.0) * (1 - p));
  const double l0 = js_log(0.0 * p + (1 - 0.0) * (1 - p));
#define JBERN_ADD(b) lp = lp + ((m >> (b)) & 1u ? l1 : l0)          // SASS: R2P per 7 points, 2 FSEL + 1 DADD per point
#pragma unroll 1
  for (int w = 0; w < N / 32; ++w) {
    const unsigned m = mask[w];
#pragma unroll
    for (int b = 0; b < 32; ++b) JBERN_ADD(b);
  }
  if (N % 32) {
    const unsigned m = mask[N / 32];
#pragma unroll
    for (int b = 0; b < N % 32; ++b) JBERN_ADD(b);
  }
#undef JBERN_ADD
  return lp;
}
__device__ __forceinline__ double jit_norm_factorised(double n, double S, double sd) { return n * (JNORM_C0 - js_log(sd)) - S / (2 * sd * sd); }

}  // namespace amwg

#include "amwg_jit_generated.inc"

namespace amwg {

extern "C" __global__ void __launch_bounds__(JTHREADS, JMINB) amwg_jit_sweep(const __grid_constant__ JitArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_res;
  const ChainArrays& a = A.a;
  const SweepArgs& sa = A.sa;
  const unsigned long long C = a.C;
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = tid < C;
  const unsigned long long chain = valid ? tid : C - 1;
  const unsigned long long gchain = a.first_chain + chain;

  if (threadIdx.x == 0) {
    mbar_init(&bar_res, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
#if JN_RES > 0
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_res, JRES_TOTAL_BYTES);
#pragma unroll
    for (int k = 0; k < JN_RES; ++k) tma_bulk_g2s(smem + JRES_OFF[k], A.col[JRES_COL[k]], JRES_BYTES[k], &bar_res);
  }
  mbar_wait(&bar_res, 0);
#endif
#if JN_BERN > 0
  for (int k = 0; k < JN_BERN; ++k) {                           // 0/1 columns -> bit masks, one ballot per 32 points
    const int n = JBERN_N[k];
    unsigned* mask = reinterpret_cast<unsigned*>(smem + JBERN_MASK[k]);
    if (threadIdx.x == 0) mask[(n + 31) / 32] = 0u;
    __syncthreads();
    for (int base = (int)(threadIdx.x & ~31u); base < n; base += JTHREADS) {
      const int i = base + (int)(threadIdx.x & 31u);
      const double yi = i < n ? lds_f64_sa(smem_u32(smem) + JBERN_DATA[k] + 8u * (unsigned)i) : 0.0;
      const unsigned ones = __ballot_sync(0xffffffffu, yi == 1.0);
      const unsigned bad = __ballot_sync(0xffffffffu, !(yi == 1.0 || yi == 0.0));
      if ((threadIdx.x & 31u) == 0) { mask[base >> 5] = ones; if (bad) atomicOr(&mask[(n + 31) / 32], 1u); }
    }
  }
  __syncthreads();
#endif
  if (!valid) return;                                           // no CTA-wide step follows: threads past the last chain are done

#if JWS_SMEM
  double* sp = reinterpret_cast<double*>(smem + JWS_OFF) + threadIdx.x;      // the chain's state: one shared-memory column per thread
  const unsigned long long ss = JTHREADS;
  for (int c = 0; c < JD; ++c) ST(c) = a.state[(unsigned long long)c * C + chain];
#else
  double* sp = a.state + chain;
  const unsigned long long ss = C;
#endif

  RandomStream g;
  g.init(a.rng_n[chain]);
  unsigned long long perm = a.perm[chain];
  double curr = a.curr_lp[chain];
#if JMAX_DIM0 > 1
  unsigned char order[JMAX_DIM0 < kLocalOrder ? JMAX_DIM0 : kLocalOrder];
#endif

  long long rec_phase = sa.record ? sa.sample_i0 % sa.thin : 0;
  long long row = sa.record ? (sa.sample_i0 + sa.thin - 1) / sa.thin : 0;
  for (long long s = 0; s < sa.n_sweeps; ++s) {
    if (sa.record) {                                            // Sampler.sample: the state BEFORE stepping (mcmc.js:1021-1027)
      const bool rec_now = rec_phase == 0;
      if (++rec_phase == sa.thin) rec_phase = 0;
      if (rec_now) {
#if JN_DERIVED > 0
        double der[JN_DERIVED];
        bool have_der = false;
#endif
        for (int j = 0; j < sa.n_monitor; ++j) {
          const int e = sa.monitor[j];
          double v;
          if (e < JD) {
            v = ST(e);
          } else {
#if JN_DERIVED > 0
            if (!have_der) { jit_derived(smem, sp, ss, der); have_der = true; }
            v = der[e - JD];
#else
            v = CUDART_NAN;
#endif
          }
          sa.out[((unsigned long long)row * sa.n_monitor + j) * C + chain] = v;
        }
        ++row;
      }
    }
    // -- AmwgStepper.step: shuffle_array(this.substeppers), in place (mcmc.js:887, 228-236)
#if JP > 1
    for (int i = JP - 1; i > 0; --i) {
      const int j = (int)floor(g.next(a.seed, gchain) * (i + 1));
      perm_swap(a, perm, chain, i, j, true);
    }
#endif
#pragma unroll 1
    for (int slot = 0; slot < JP; ++slot) {
      const int p = (JP > 1) ? perm_get(a, perm, chain, slot) : 0;
      const int n_comp = jp_ncomp(p), off = jp_off(p), ptype = jp_type(p);
      const double lower = jp_lower(p), upper = jp_upper(p);
#if JMAX_DIM0 > 1
      const int dim0 = jp_dim0(p);
      const int inner = n_comp / dim0;
      if (n_comp > 1) {                                         // nested_array_random_apply: top level only (mcmc.js:246-252)
        for (int i = 0; i < dim0; ++i) ord_set(a, order, chain, dim0, i, i, true);
        for (int i = dim0 - 1; i > 0; --i) {
          const int j = (int)floor(g.next(a.seed, gchain) * (i + 1));
          const int t = ord_get(a, order, chain, dim0, i);
          ord_set(a, order, chain, dim0, i, ord_get(a, order, chain, dim0, j), true);
          ord_set(a, order, chain, dim0, j, t, true);
        }
      }
#endif
#pragma unroll 1
      for (int r = 0; r < n_comp; ++r) {
        int c = off;
#if JMAX_DIM0 > 1
        if (n_comp > 1) c += ord_get(a, order, chain, dim0, r / inner) * inner + (r % inner);
#endif
        const double cur = ST(c);
        if (ptype == AMWG_BINARY) {
          // BinaryStepper.step (mcmc.js:753-767); log_post of the current value is the cached one
          const double other = (cur == 0.0) ? 1.0 : 0.0;
          const double lp_new = jit_logpost(smem, sp, ss, c, other);
          const double z0raw = (cur == 0.0) ? curr : lp_new, z1raw = (cur == 0.0) ? lp_new : curr;
          const double mx = js_max(z0raw, z1raw);
          const double z0 = z0raw - mx, z1 = z1raw - mx;
          const double zero_prob = js_exp(z0 - js_log(js_exp(z0) + js_exp(z1)));
          const bool zero = g.next(a.seed, gchain) < zero_prob;
          ST(c) = zero ? 0.0 : 1.0;
          curr = zero ? z0raw : z1raw;
          continue;
        }
        // generate_proposal (mcmc.js:519, 577-579 / 596-598) and the bounds check (:520)
        double prop = js_rnorm(g, a.seed, gchain, cur, a.psd[(unsigned long long)c * C + chain]);
        if (ptype == AMWG_INT) prop = js_round(prop);
        if (prop < lower || prop > upper) continue;             // rejected without evaluation, no accept uniform
        const double lp_new = jit_logpost(smem, sp, ss, c, prop);
        const double accept_prob = js_exp(lp_new - curr);       // Metropolis accept (mcmc.js:527-534): strict >, NaN rejects
        if (accept_prob > g.next(a.seed, gchain)) {
          curr = lp_new;
          ST(c) = prop;
          if (A.adapting[c]) a.acc[(unsigned long long)c * C + chain] += 1;
        }
      }
    }
  }
  a.rng_n[chain] = g.n;
  a.perm[chain] = perm;
  a.curr_lp[chain] = curr;
#if JWS_SMEM
  for (int c = 0; c < JD; ++c) a.state[(unsigned long long)c * C + chain] = ST(c);
#endif
}

}  // namespace amwg
