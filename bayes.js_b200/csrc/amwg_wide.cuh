// amwg_wide.cuh -- EXPERIMENTAL variant of the sweep kernel with W chains per thread (off by default, AMWG_CHAINS_PER_THREAD=2|4).
// Bit-exact with the default kernel (the GPU test-suite passes with W = 2 and 4) but slower on B200: see DESIGN.md section 4.
//
// Same semantics as amwg_sweep_kernel (every chain is still an independent run of the reference under its own Philox stream);
// what changes is the mapping: thread t owns chains t, t+T, ... (T = threads), and walks them together, so that
//   * the plate loop reads each data point once per thread and feeds W chains (4 LDS.128 + W*16 fp64 instructions per 8 points),
//   * the interpreter fetches / decodes / dispatches each program word once for W chains,
//   * W independent dependency chains per thread hide the latency of the per-step bookkeeping (Philox, log/exp, loads).
// All per-chain quantities are W-element arrays in registers; control flow is the union over the thread's chains with per-chain
// predicates (a chain whose parameter has fewer components, or whose proposal is out of bounds, simply does not commit).
#pragma once

namespace amwg {

template <int W>
struct WideEval {
  const double* st[W];        // state base + chain
  unsigned long long stride;
  int moved[W];
  double val[W];
  __device__ __forceinline__ double comp(int k, int c) const { return c == moved[k] ? val[k] : st[k][(unsigned long long)c * stride]; }
};

// sum_i (x_i - mean_k)^2 for W chains at once: each 16-byte broadcast load feeds W DADD+DFMA pairs per point.
template <int W>
__device__ __forceinline__ void sum_sq_dev_w(const double* __restrict__ x, unsigned saddr, int n, const double (&mean)[W], double (&S)[W]) {
  double s0[W], s1[W], s2[W], s3[W];
#pragma unroll
  for (int k = 0; k < W; ++k) { s0[k] = 0.0; s1[k] = 0.0; s2[k] = 0.0; s3[k] = 0.0; }
  int i = 0;
  if ((reinterpret_cast<unsigned long long>(x) & 15ull) && n > 0) {
    double xi = x[0];
#pragma unroll
    for (int k = 0; k < W; ++k) { double d = xi - mean[k]; s3[k] = fma(d, d, s3[k]); }
    i = 1;
  }
#define AMWG_ACCW(P0, P1, P2, P3)                                                             \
  _Pragma("unroll") for (int k = 0; k < W; ++k) {                                             \
    double d0 = P0.x - mean[k], d1 = P0.y - mean[k], d2 = P1.x - mean[k], d3 = P1.y - mean[k]; \
    double d4 = P2.x - mean[k], d5 = P2.y - mean[k], d6 = P3.x - mean[k], d7 = P3.y - mean[k]; \
    s0[k] = fma(d0, d0, s0[k]); s1[k] = fma(d1, d1, s1[k]); s2[k] = fma(d2, d2, s2[k]); s3[k] = fma(d3, d3, s3[k]); \
    s0[k] = fma(d4, d4, s0[k]); s1[k] = fma(d5, d5, s1[k]); s2[k] = fma(d6, d6, s2[k]); s3[k] = fma(d7, d7, s3[k]); \
  }
  const int nb = (n - i) >> 3;
  if (nb > 0) {
    if (saddr) {
      unsigned a = saddr + 8u * (unsigned)i;
      double2 p0 = lds_f64x2(a), p1 = lds_f64x2(a + 16u), p2 = lds_f64x2(a + 32u), p3 = lds_f64x2(a + 48u);
#pragma unroll 2
      for (int b = 1; b < nb; ++b) {
        a += 64u;
        double2 q0 = lds_f64x2(a), q1 = lds_f64x2(a + 16u), q2 = lds_f64x2(a + 32u), q3 = lds_f64x2(a + 48u);
        AMWG_ACCW(p0, p1, p2, p3)
        p0 = q0; p1 = q1; p2 = q2; p3 = q3;
      }
      AMWG_ACCW(p0, p1, p2, p3)
    } else {
      const double2* g = reinterpret_cast<const double2*>(x + i);
#pragma unroll 2
      for (int b = 0; b < nb; ++b, g += 4) {
        double2 p0 = g[0], p1 = g[1], p2 = g[2], p3 = g[3];
        AMWG_ACCW(p0, p1, p2, p3)
      }
    }
    i += nb << 3;
  }
#undef AMWG_ACCW
  for (; i < n; ++i) {
    double xi = x[i];
#pragma unroll
    for (int k = 0; k < W; ++k) { double d = xi - mean[k]; s0[k] = fma(d, d, s0[k]); }
  }
#pragma unroll
  for (int k = 0; k < W; ++k) S[k] = (s0[k] + s1[k]) + (s2[k] + s3[k]);
}

template <int W>
__device__ __noinline__ void plate_norm_iid_w(const Ctx& ctx, int q, const double (&mean)[W], const double (&sd)[W], double (&out)[W]) {
  const amwg_plate& pl = ctx.plates[q];
  int c = pl.col[0], off = pl.iparam[2];
  unsigned sa = ctx.col_saddr[c] ? ctx.col_saddr[c] + 8u * (unsigned)off : 0u;
  double S[W];
  sum_sq_dev_w<W>(ctx.col[c] + off, sa, pl.n, mean, S);
#pragma unroll
  for (int k = 0; k < W; ++k) out[k] = norm_factorised(ctx, (double)pl.n, S[k], sd[k]);
}

template <int W>
__device__ __noinline__ void plate_bern_iid_w(const Ctx& ctx, int q, const double (&p)[W], double (&lp)[W]) {
  const amwg_plate& pl = ctx.plates[q];
  double l1[W], l0[W];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    l1[k] = js_log(1.0 * p[k] + (1 - 1.0) * (1 - p[k]));
    l0[k] = js_log(0.0 * p[k] + (1 - 0.0) * (1 - p[k]));
  }
  const double* __restrict__ y = ctx.col[pl.col[0]] + pl.iparam[2];
  for (int i = 0; i < pl.n; ++i) {
    double yi = y[i];
#pragma unroll
    for (int k = 0; k < W; ++k) lp[k] = lp[k] + (yi == 1.0 ? l1[k] : (yi == 0.0 ? l0[k] : -CUDART_INF));
  }
}

template <int W>
__device__ __noinline__ void plate_norm_grouped_w(const Ctx& ctx, int q, const WideEval<W>& es, const double (&sd)[W], double (&out)[W]) {
  const amwg_plate& pl = ctx.plates[q];
  int c = pl.col[0], off = pl.iparam[2];
  const double* __restrict__ start = ctx.col[pl.col[1]];
  int J = pl.iparam[1], base = pl.iparam[0];
  double S[W];
#pragma unroll
  for (int k = 0; k < W; ++k) S[k] = 0.0;
  for (int j = 0; j < J; ++j) {
    int a = (int)start[j] + off, b = (int)start[j + 1] + off;
    unsigned sa = ctx.col_saddr[c] ? ctx.col_saddr[c] + 8u * (unsigned)a : 0u;
    double mean[W], Sj[W];
#pragma unroll
    for (int k = 0; k < W; ++k) mean[k] = es.comp(k, base + j);
    sum_sq_dev_w<W>(ctx.col[c] + a, sa, b - a, mean, Sj);
#pragma unroll
    for (int k = 0; k < W; ++k) S[k] = S[k] + Sj[k];
  }
#pragma unroll
  for (int k = 0; k < W; ++k) out[k] = norm_factorised(ctx, (double)pl.n, S[k], sd[k]);
}

// log_post for the W chains of a thread: one fetch/decode/dispatch per program word, W arithmetic results.
template <int W>
__device__ __noinline__ void run_logpost_w(unsigned code_sa, unsigned consts_sa, const Ctx& ctx, const WideEval<W>& es, int pc, double (&lp_out)[W]) {
  double stk[kStack][W];
  double tos[W], lp[W];
#pragma unroll
  for (int k = 0; k < W; ++k) { tos[k] = 0.0; lp[k] = 0.0; }
  int sp = 0;
  int loop_i = 0, loop_n = 0;
#define WNEXT() ((int)lds_u32(code_sa + 4u * (unsigned)(pc++)))
#define WPOP(dst) do { --sp; _Pragma("unroll") for (int k = 0; k < W; ++k) { dst[k] = tos[k]; tos[k] = stk[sp][k]; } } while (0)
#define WOPND(dst, mode)                                                                                      \
  do {                                                                                                        \
    if ((mode) == AMWG_MODE_STACK) { WPOP(dst); }                                                             \
    else {                                                                                                    \
      int _ix = WNEXT();                                                                                      \
      if ((mode) == AMWG_MODE_CONST) { double _c = lds_f64(consts_sa + 8u * (unsigned)_ix); _Pragma("unroll") for (int k = 0; k < W; ++k) dst[k] = _c; } \
      else { _Pragma("unroll") for (int k = 0; k < W; ++k) dst[k] = es.comp(k, _ix); }                          \
    }                                                                                                         \
  } while (0)
#define WEACH(expr) _Pragma("unroll") for (int k = 0; k < W; ++k) { r[k] = (expr); } break;
  for (;;) {
    const unsigned w = (unsigned)WNEXT();
    const int op = w & 0xff;
    const bool acc = (w >> 16) & 1;
    const bool store = (w & AMWG_STORE_FLAG) != 0;      // term-cache ids in the full program are skipped: this variant does not cache
    const int a = (int)(w >> 18);
    double x[W], y[W], z[W], t[W], r[W];
#pragma unroll
    for (int k = 0; k < W; ++k) { x[k] = 0.0; y[k] = 0.0; z[k] = 0.0; t[k] = 0.0; r[k] = 0.0; }
    if (op != AMWG_OP_PLATE) {
      const int mD = (w >> 14) & 3, mC = (w >> 12) & 3, mB = (w >> 10) & 3, mA = (w >> 8) & 3;
      if (mD != AMWG_MODE_NONE) WOPND(t, mD);
      if (mC != AMWG_MODE_NONE) WOPND(z, mC);
      if (mB != AMWG_MODE_NONE) WOPND(y, mB);
      if (mA != AMWG_MODE_NONE) WOPND(x, mA);
    }
    bool has_r = true;
    switch (op) {
      case AMWG_OP_CONST: { double c = lds_f64(consts_sa + 8u * (unsigned)a); WEACH(c) }
      case AMWG_OP_COMP: WEACH(es.comp(k, a))
      case AMWG_OP_DATA: { double c = ctx.col[a][WNEXT()]; WEACH(c) }
      case AMWG_OP_DATA_I: { int off = WNEXT(); int stride = WNEXT(); double c = ctx.col[a][off + stride * loop_i]; WEACH(c) }
      case AMWG_OP_COMP_I: {
        int off = WNEXT(); int stride = WNEXT(); int base = WNEXT();
        int ci = base + (int)ctx.col[a][off + stride * loop_i];
        WEACH(es.comp(k, ci))
      }
      case AMWG_OP_ADD: WEACH(x[k] + y[k])
      case AMWG_OP_SUB: WEACH(x[k] - y[k])
      case AMWG_OP_MUL: WEACH(x[k] * y[k])
      case AMWG_OP_DIV: WEACH(x[k] / y[k])
      case AMWG_OP_NEG: WEACH(-x[k])
      case AMWG_OP_LOG: WEACH(js_log(x[k]))
      case AMWG_OP_EXP: WEACH(js_exp(x[k]))
      case AMWG_OP_SQRT: WEACH(sqrt(x[k]))
      case AMWG_OP_ABS: WEACH(fabs(x[k]))
      case AMWG_OP_POW: WEACH((y[k] == 2.0) ? x[k] * x[k] : cold_op(op, x[k], y[k], z[k], t[k]))
      case AMWG_OP_LT: WEACH(x[k] < y[k] ? 1.0 : 0.0)
      case AMWG_OP_LE: WEACH(x[k] <= y[k] ? 1.0 : 0.0)
      case AMWG_OP_GT: WEACH(x[k] > y[k] ? 1.0 : 0.0)
      case AMWG_OP_GE: WEACH(x[k] >= y[k] ? 1.0 : 0.0)
      case AMWG_OP_EQ: WEACH(x[k] == y[k] ? 1.0 : 0.0)
      case AMWG_OP_NE: WEACH(x[k] != y[k] ? 1.0 : 0.0)
      case AMWG_OP_AND: WEACH((x[k] != 0.0 && y[k] != 0.0) ? 1.0 : 0.0)
      case AMWG_OP_OR: WEACH((x[k] != 0.0 || y[k] != 0.0) ? 1.0 : 0.0)
      case AMWG_OP_NOT: WEACH(x[k] != 0.0 ? 0.0 : 1.0)
      case AMWG_OP_SELECT: WEACH(x[k] != 0.0 ? y[k] : z[k])
      case AMWG_OP_NORM_K: WEACH(z[k] - ((x[k] - y[k]) * (x[k] - y[k])) / t[k])
      case AMWG_OP_UNIF_K: WEACH((x[k] < y[k] || x[k] > z[k]) ? -CUDART_INF : t[k])
      case AMWG_OP_BETA_K: WEACH((x[k] > 1 || x[k] < 0) ? -CUDART_INF : (y[k] * js_log(x[k]) + z[k] * js_log(1 - x[k])) - t[k])
      case AMWG_OP_ACC: {
        double v[W]; WPOP(v);
#pragma unroll
        for (int k = 0; k < W; ++k) lp[k] = lp[k] + v[k];
        has_r = false;
        break;
      }
      case AMWG_OP_PLATE: {
        has_r = false;
        const int kind = ctx.plates[a].kind;
        if (kind == AMWG_PLATE_NORM_IID) {
          double mean[W], sd[W], v[W];
          WOPND(sd, (w >> 10) & 3); WOPND(mean, (w >> 8) & 3);
          plate_norm_iid_w<W>(ctx, a, mean, sd, v);
#pragma unroll
          for (int k = 0; k < W; ++k) lp[k] = lp[k] + v[k];
        } else if (kind == AMWG_PLATE_BERN_IID) {
          double p[W]; WOPND(p, (w >> 8) & 3);
          plate_bern_iid_w<W>(ctx, a, p, lp);
        } else if (kind == AMWG_PLATE_NORM_GROUPED) {
          double sd[W], v[W]; WOPND(sd, (w >> 8) & 3);
          plate_norm_grouped_w<W>(ctx, a, es, sd, v);
#pragma unroll
          for (int k = 0; k < W; ++k) lp[k] = lp[k] + v[k];
        } else if (kind == AMWG_PLATE_POIS_LOGLIN) {
#pragma unroll
          for (int k = 0; k < W; ++k) {
            EvalState e1{es.st[k], es.stride, es.moved[k], es.val[k]};
            lp[k] = lp[k] + plate_pois_loglin(ctx, a, e1);
          }
        }
        if (store) (void)WNEXT();
        break;
      }
      case AMWG_OP_LOOP_BEGIN: {
        int skip_to = WNEXT();
        loop_i = 0; loop_n = ctx.plates[a].n;
        if (loop_n <= 0) pc = skip_to;
        has_r = false;
        break;
      }
      case AMWG_OP_LOOP_END: {
        int body = WNEXT();
        double v[W]; WPOP(v);
#pragma unroll
        for (int k = 0; k < W; ++k) lp[k] = lp[k] + v[k];
        if (++loop_i < loop_n) pc = body; else loop_i = 0;
        has_r = false;
        break;
      }
      case AMWG_OP_END:
#pragma unroll
        for (int k = 0; k < W; ++k) lp_out[k] = lp[k];
        return;
      default: WEACH(cold_op(op, x[k], y[k], z[k], t[k]))
    }
    if (has_r) {
      if (acc) {
#pragma unroll
        for (int k = 0; k < W; ++k) lp[k] = lp[k] + r[k];
        if (store) (void)WNEXT();
      } else {
#pragma unroll
        for (int k = 0; k < W; ++k) { stk[sp][k] = tos[k]; tos[k] = r[k]; }
        ++sp;
      }
    }
  }
#undef WNEXT
#undef WPOP
#undef WOPND
#undef WEACH
}

#ifndef AMWG_WIDE_MINBLOCKS
#define AMWG_WIDE_MINBLOCKS 3
#endif

template <int W>
__global__ void __launch_bounds__(kThreads, AMWG_WIDE_MINBLOCKS) amwg_sweep_kernel_wide(ModelDev m, ChainArrays a, SweepArgs sa) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ Ctx ctx;
  __shared__ __align__(8) unsigned long long bar;
  stage_model(m, smem, ctx, &bar);
  if (threadIdx.x == 0) ctx.ring_saddr = 0u;          // this variant does not run CTA-uniformly: columns outside shared memory come from L2
  __syncthreads();

  const unsigned long long C = a.C;
  const unsigned long long T = (C + W - 1) / W;                      // threads; thread t owns chains t, t+T, ...
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= T) return;
  unsigned long long chain[W];
  bool valid[W];
  RandomStream g[W];
  unsigned long long perm[W];
  double curr[W];
  WideEval<W> es;
  es.stride = C;
#pragma unroll
  for (int k = 0; k < W; ++k) {
    unsigned long long ch = tid + (unsigned long long)k * T;
    valid[k] = ch < C;
    chain[k] = valid[k] ? ch : C - 1;                                // a missing chain shadows the last one and writes nothing
    g[k].init(a.rng_n[chain[k]]);
    perm[k] = a.perm[chain[k]];
    curr[k] = a.curr_lp[chain[k]];
    es.st[k] = a.state + chain[k];
  }
  const int P = m.n_params;
  const unsigned code_sa = smem_u32(ctx.code), consts_sa = smem_u32(ctx.consts);
  unsigned char order[W][kMaxDim0];

  long long rec_phase = sa.record ? sa.sample_i0 % sa.thin : 0;
  long long row = sa.record ? (sa.sample_i0 + sa.thin - 1) / sa.thin : 0;
  for (long long s = 0; s < sa.n_sweeps; ++s) {
    // -- Sampler.sample: record the state BEFORE stepping (mcmc.js:1021-1027)
    if (sa.record) {
      const bool rec_now = rec_phase == 0;
      if (++rec_phase == sa.thin) rec_phase = 0;
      if (rec_now) {
#pragma unroll
        for (int k = 0; k < W; ++k) {
          if (!valid[k]) continue;
          double der[kMaxDerived];
          bool have_der = false;
          for (int j = 0; j < sa.n_monitor; ++j) {
            int e = sa.monitor[j];
            double v;
            if (e < m.D) {
              v = es.st[k][(unsigned long long)e * C];
            } else {
              if (!have_der) { EvalState e1{es.st[k], C, -1, 0.0}; run_ctx(ctx, e1, derived_pc(m, e1), der, false); have_der = true; }
              v = der[e - m.D];
            }
            sa.out[((unsigned long long)row * sa.n_monitor + j) * C + chain[k]] = v;
          }
        }
        ++row;
      }
    }
    // -- AmwgStepper.step: shuffle_array(this.substeppers), in place (mcmc.js:887, 228-236)
    for (int i = P - 1; i > 0; --i) {
#pragma unroll
      for (int k = 0; k < W; ++k) {
        int j = (int)floor(g[k].next(a.seed, a.first_chain + chain[k]) * (i + 1));
        unsigned long long vi = (perm[k] >> (4 * i)) & 15ull, vj = (perm[k] >> (4 * j)) & 15ull;
        perm[k] = (perm[k] & ~(15ull << (4 * i))) | (vj << (4 * i));
        perm[k] = (perm[k] & ~(15ull << (4 * j))) | (vi << (4 * j));
      }
    }
    for (int slot = 0; slot < P; ++slot) {
      int pidx[W], n_rounds[W], R = 0;
#pragma unroll
      for (int k = 0; k < W; ++k) {
        pidx[k] = (int)((perm[k] >> (4 * slot)) & 15ull);
        const amwg_param& pa = ctx.params[pidx[k]];
        n_rounds[k] = pa.n_comp;
        R = max(R, n_rounds[k]);
        if (pa.n_comp > 1) {
          // nested_array_random_apply: fresh identity, shuffled, top level only (mcmc.js:246-252)
          for (int i = 0; i < pa.dim0; ++i) order[k][i] = (unsigned char)i;
          for (int i = pa.dim0 - 1; i > 0; --i) {
            int j = (int)floor(g[k].next(a.seed, a.first_chain + chain[k]) * (i + 1));
            unsigned char t = order[k][i]; order[k][i] = order[k][j]; order[k][j] = t;
          }
        }
      }
      for (int r = 0; r < R; ++r) {
        int c[W];
        double cur[W], prop[W];
        bool active[W], need[W], binary[W];
        // ---- propose (mcmc.js:519-522, 577-579 / 596-598; binary: the state value whose log_post is not cached)
#pragma unroll
        for (int k = 0; k < W; ++k) {
          const amwg_param& pa = ctx.params[pidx[k]];
          active[k] = r < n_rounds[k];
          const int rr = active[k] ? r : 0;
          const int inner = pa.n_comp / pa.dim0;
          c[k] = pa.comp_offset + (pa.n_comp > 1 ? (int)order[k][rr / inner] * inner + (rr % inner) : 0);
          cur[k] = es.st[k][(unsigned long long)c[k] * C];
          binary[k] = pa.type == AMWG_BINARY;
          prop[k] = cur[k];
          need[k] = false;
          if (active[k]) {
            if (binary[k]) {
              prop[k] = (cur[k] == 0.0) ? 1.0 : 0.0;
              need[k] = true;
            } else {
              prop[k] = js_rnorm(g[k], a.seed, a.first_chain + chain[k], cur[k], a.psd[(unsigned long long)c[k] * C + chain[k]]);
              if (pa.type == AMWG_INT) prop[k] = js_round(prop[k]);
              need[k] = !(prop[k] < pa.lower || prop[k] > pa.upper);
            }
          }
          es.moved[k] = c[k];
          es.val[k] = prop[k];
        }
        // ---- evaluate log_post at the proposals: one walk over the program and the data for the W chains
        bool any = false;
#pragma unroll
        for (int k = 0; k < W; ++k) any = any || need[k];
        double lp_new[W];
#pragma unroll
        for (int k = 0; k < W; ++k) lp_new[k] = 0.0;
        __syncwarp(__activemask());
        if (any) {
          if (m.n_variant_comps == 0) {
            run_logpost_w<W>(code_sa, consts_sa, ctx, es, m.logpost_prog, lp_new);
          } else {                                   // program depends on each chain's binary configuration: evaluate chain by chain
#pragma unroll
            for (int k = 0; k < W; ++k) {
              EvalStateT<false> e1{es.st[k], C, es.moved[k], es.val[k]};
              lp_new[k] = eval_logpost<false>(ctx, e1, logpost_pc(m, e1));
            }
          }
        }
        // ---- accept / reject
#pragma unroll
        for (int k = 0; k < W; ++k) {
          if (!active[k]) continue;
          const unsigned long long ci = (unsigned long long)c[k] * C + chain[k];
          if (binary[k]) {
            // BinaryStepper.step (mcmc.js:753-767); log_post of the current value is the cached one
            double z0raw = (cur[k] == 0.0) ? curr[k] : lp_new[k], z1raw = (cur[k] == 0.0) ? lp_new[k] : curr[k];
            double mx = js_max(z0raw, z1raw);
            double z0 = z0raw - mx, z1 = z1raw - mx;
            double zero_prob = js_exp(z0 - js_log(js_exp(z0) + js_exp(z1)));
            bool zero = g[k].next(a.seed, a.first_chain + chain[k]) < zero_prob;
            if (valid[k]) a.state[ci] = zero ? 0.0 : 1.0;
            curr[k] = zero ? z0raw : z1raw;
          } else if (need[k]) {
            // Metropolis accept (mcmc.js:527-534): strict >, NaN rejects
            double accept_prob = js_exp(lp_new[k] - curr[k]);
            if (accept_prob > g[k].next(a.seed, a.first_chain + chain[k])) {
              curr[k] = lp_new[k];
              if (valid[k]) {
                a.state[ci] = prop[k];
                if (m.adapting[c[k]]) a.acc[ci] += 1;
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < W; ++k) {
    if (!valid[k]) continue;
    a.rng_n[chain[k]] = g[k].n;
    a.perm[chain[k]] = perm[k];
    a.curr_lp[chain[k]] = curr[k];
  }
}

}  // namespace amwg
