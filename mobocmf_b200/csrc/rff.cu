// Pareto-sample generation on the GPU: evaluation of random-Fourier-feature function samples
// of an MFDGP layer chain on a grid of candidate points, and the non-dominated cull of the objective values.
//
// Replaces the numpy evaluation of the closures returned by
//   MFDGPHiddenLayer._sample_from_posterior(_layer0) / _sample_from_prior(_layer0)
//                                              (mobocmf/layers/mfdgp_hidden_layer.py:288-293, 311-514: `wrapper`)
// when MOOP.compute_pareto_solution_from_samples evaluates them on its 1000 d^2-point grid
//                                              (mobocmf/util/moop.py:221-286),
// and MOOP.compute_pareto_front / obtain_indices_pareto (mobocmf/util/moop.py:141-185).
//
// Function sample of layer 0:   f_0(x) = sum_j theta_j s cos(W_j . x + b_j),               s = sqrt(2 alpha / F)
// layer l >= 1, with f = f_{l-1}(x) (the chain: the lower fidelity's sample is an input of the higher one):
//   f_l(x) = sum_j th1_j s1 cos(a1_j) f + th2_j s1f cos(a1_j + Wf_j f) + th3_j s2 cos(a2_j)
//   a1_j = Wx1_j . x + bx1_j,  a2_j = Wx2_j . x + bx2_j,  s1 = sqrt(2 alpha_x1 / F) sqrt(nu_lin), ...
// and its x-gradient by the chain rule through f.  One CTA: 32 grid points (lane <-> point) x 8 feature slices
// (warp <-> features j = warp, warp + 8, ...); W_j, b_j, theta_j are warp-uniform (broadcast) loads, per-point partial
// sums are folded over the 8 warps in fixed order (deterministic), layers run in sequence inside the CTA.
// FP64 (DFMA + the cos / sincos polynomial) bound: ~60 FP64 operations per (point, feature).
#include "common.cuh"

namespace mobo {

constexpr int RFF_THREADS = 256, RFF_WARPS = RFF_THREADS / 32, RFF_PTS = 32, RFF_MAX_LAYERS = 4;

// parameter block of a layer's function sample (device doubles):
//   layer 0 : [W (F x d) | b (F) | theta (F)]
//   layer>=1: [Wx1 (F x d) | Wf (F) | Wx2 (F x d) | bx1 (F) | bx2 (F) | theta (3 F)]
struct RffArgs {
  int L, d, F, want_grad;
  const double* params[RFF_MAX_LAYERS];
  double scale[RFF_MAX_LAYERS][3];     // layer 0: {s, -, -};  layer >= 1: {s1 (with sqrt(nu_lin)), s1f, s2}
  const double* x;                     // n x d
  long long n;
  double* f;                           // L x n : every layer's value
  double* grad;                        // n x d : d f_{L-1} / d x  (optional)
};

template <int D, bool GRAD>
__global__ void __launch_bounds__(RFF_THREADS) rff_eval_kernel(const __grid_constant__ RffArgs a) {
  __shared__ double part[RFF_WARPS][1 + D][RFF_PTS];
  __shared__ double fprev[RFF_PTS], dfprev[D][RFF_PTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long pt = (long long)blockIdx.x * RFF_PTS + lane;
  const bool ok = pt < a.n;
  const int F = a.F;
  double x[D];
#pragma unroll
  for (int c = 0; c < D; ++c) x[c] = ok ? a.x[(size_t)pt * D + c] : 0.0;

  for (int l = 0; l < a.L; ++l) {
    const double* P = a.params[l];
    double acc = 0.0, g[D];
#pragma unroll
    for (int c = 0; c < D; ++c) g[c] = 0.0;
    if (l == 0) {
      const double* W = P;
      const double* b = P + (size_t)F * D;
      const double* th = b + F;
      const double s = a.scale[0][0];
      for (int j = warp; j < F; j += RFF_WARPS) {
        double arg = __ldg(b + j), w[D];
#pragma unroll
        for (int c = 0; c < D; ++c) { w[c] = __ldg(W + (size_t)j * D + c); arg = fma(w[c], x[c], arg); }
        const double ts = __ldg(th + j) * s;
        if (GRAD) {
          double sn, cs;
          sincos(arg, &sn, &cs);
          acc = fma(ts, cs, acc);
          const double q = -ts * sn;
#pragma unroll
          for (int c = 0; c < D; ++c) g[c] = fma(q, w[c], g[c]);
        } else {
          acc = fma(ts, cos(arg), acc);
        }
      }
    } else {
      const double* Wx1 = P;
      const double* Wf = Wx1 + (size_t)F * D;
      const double* Wx2 = Wf + F;
      const double* bx1 = Wx2 + (size_t)F * D;
      const double* bx2 = bx1 + F;
      const double* th = bx2 + F;
      const double s1 = a.scale[l][0], s1f = a.scale[l][1], s2 = a.scale[l][2];
      const double f = fprev[lane];
      double df[D];
#pragma unroll
      for (int c = 0; c < D; ++c) df[c] = GRAD ? dfprev[c][lane] : 0.0;
      for (int j = warp; j < F; j += RFF_WARPS) {
        double a1 = __ldg(bx1 + j), a2 = __ldg(bx2 + j), w1[D], w2[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
          w1[c] = __ldg(Wx1 + (size_t)j * D + c);
          w2[c] = __ldg(Wx2 + (size_t)j * D + c);
          a1 = fma(w1[c], x[c], a1);
          a2 = fma(w2[c], x[c], a2);
        }
        const double wf = __ldg(Wf + j);
        const double a1f = fma(wf, f, a1);
        const double t1 = __ldg(th + j) * s1, t2 = __ldg(th + F + j) * s1f, t3 = __ldg(th + 2 * F + j) * s2;
        if (GRAD) {
          double sn1, cs1, sn1f, cs1f, sn2, cs2;
          sincos(a1, &sn1, &cs1);
          sincos(a1f, &sn1f, &cs1f);
          sincos(a2, &sn2, &cs2);
          acc = fma(t1 * cs1, f, acc);
          acc = fma(t2, cs1f, acc);
          acc = fma(t3, cs2, acc);
          const double q1 = -t1 * sn1 * f, q1c = t1 * cs1, q1f = -t2 * sn1f, q2 = -t3 * sn2;
#pragma unroll
          for (int c = 0; c < D; ++c) {
            double gc = g[c];
            gc = fma(q1, w1[c], gc);
            gc = fma(q1c, df[c], gc);
            gc = fma(q1f, fma(wf, df[c], w1[c]), gc);
            gc = fma(q2, w2[c], gc);
            g[c] = gc;
          }
        } else {
          acc = fma(t1 * cos(a1), f, acc);
          acc = fma(t2, cos(a1f), acc);
          acc = fma(t3, cos(a2), acc);
        }
      }
    }
    part[warp][0][lane] = acc;
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < D; ++c) part[warp][1 + c][lane] = g[c];
    }
    __syncthreads();
    // fold the feature slices in fixed order: value by warp 0, gradient component c by warp 1 + c (mod 8)
    if (warp == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < RFF_WARPS; ++w) s += part[w][0][lane];
      fprev[lane] = s;
      if (ok) a.f[(size_t)l * a.n + pt] = s;
    }
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < D; ++c)
        if (warp == (1 + c) % RFF_WARPS) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < RFF_WARPS; ++w) s += part[w][1 + c][lane];
          dfprev[c][lane] = s;
          if (ok && l == a.L - 1) a.grad[(size_t)pt * D + c] = s;
        }
    }
    __syncthreads();
  }
}

template <int D>
static void launch_rff_d(const RffArgs& a, cudaStream_t st) {
  const int grid = (int)((a.n + RFF_PTS - 1) / RFF_PTS);
  if (a.want_grad) MOBO_LAUNCH("rff_eval_kernel<grad>", st, rff_eval_kernel<D, true><<<grid, RFF_THREADS, 0, st>>>(a));
  else MOBO_LAUNCH("rff_eval_kernel", st, rff_eval_kernel<D, false><<<grid, RFF_THREADS, 0, st>>>(a));
}

int launch_rff_eval(const RffArgs& a, cudaStream_t st) {
  if (a.L < 1 || a.L > RFF_MAX_LAYERS || a.d < 1 || a.d > kMaxD || a.F < 1) return -2;
  if (a.n <= 0) return 0;
  switch (a.d) {
    case 1: launch_rff_d<1>(a, st); break;
    case 2: launch_rff_d<2>(a, st); break;
    case 3: launch_rff_d<3>(a, st); break;
    case 4: launch_rff_d<4>(a, st); break;
    case 5: launch_rff_d<5>(a, st); break;
    case 6: launch_rff_d<6>(a, st); break;
    case 7: launch_rff_d<7>(a, st); break;
    default: launch_rff_d<8>(a, st); break;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------
// Non-dominated cull of n points with k objectives (minimisation): mask[j] = 1 unless some point i is nowhere larger
// than j and differs from it (or equals it and has the smaller index: duplicates keep their first copy; the
// reference keeps the first copy in ITS visiting order, util/moop.py:170-185, the surviving VALUES are identical).
// One thread per candidate j, the points stream through shared memory in tiles.  Compare-bound, n^2 k / 2 on average
// (a dominated thread stops reading).
// ---------------------------------------------------------------------------------------------------
constexpr int PM_THREADS = 256, PM_MAX_K = 8;

__global__ void __launch_bounds__(PM_THREADS) pareto_mask_kernel(const double* __restrict__ pts, long long n, int k,
                                                                 unsigned char* __restrict__ mask) {
  __shared__ double tile[PM_THREADS * PM_MAX_K];
  const long long j = (long long)blockIdx.x * PM_THREADS + threadIdx.x;
  double me[PM_MAX_K];
  for (int c = 0; c < PM_MAX_K; ++c) me[c] = (j < n && c < k) ? pts[(size_t)j * k + c] : 0.0;
  bool dominated = false;
  for (long long i0 = 0; i0 < n; i0 += PM_THREADS) {
    const int cnt = (int)min((long long)PM_THREADS, n - i0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < cnt * k; idx += PM_THREADS) tile[idx] = pts[(size_t)i0 * k + idx];
    __syncthreads();
    if (j < n && !dominated) {
      for (int ii = 0; ii < cnt; ++ii) {
        bool all_le = true, any_lt = false;
        for (int c = 0; c < k; ++c) {
          const double v = tile[ii * k + c];
          all_le = all_le && (v <= me[c]);
          any_lt = any_lt || (v < me[c]);
        }
        if (all_le && (any_lt || i0 + ii < j)) { dominated = true; break; }
      }
    }
  }
  if (j < n) mask[j] = dominated ? 0 : 1;
}

int launch_pareto_mask(const double* pts, long long n, int k, unsigned char* mask, cudaStream_t st) {
  if (k < 1 || k > PM_MAX_K) return -2;
  if (n <= 0) return 0;
  MOBO_LAUNCH("pareto_mask_kernel", st,
              pareto_mask_kernel<<<(int)((n + PM_THREADS - 1) / PM_THREADS), PM_THREADS, 0, st>>>(pts, n, k, mask));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mobo
