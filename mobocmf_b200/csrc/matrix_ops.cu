// Per-step M x M operator chain of the sparse-GP layers (forward and backward) for sm_100a, batched over layers.
//
// Forward (what the reference reaches through UnwhitenedVariationalStrategy.forward / kl_mvn_mvn [upstream gpytorch],
// called from mobocmf/layers/mfdgp_hidden_layer.py:286 and mobocmf/mlls/variational_elbo_mf.py:40):
//   P = K(Z_l, Z_l) + jitter I      (CovarianceMatrixMF.add_jitter, layers/mfdgp_hidden_layer.py:17-20)
//   L = chol(P);  W = L^-1;  H = W tril(L_q);  beta = W m
//   KL = 1/2 [ 2 sum log L_ii - sum log L_q,ii^2 + |beta|^2 + |H|_F^2 - M ]
// Backward: the loss depends on P only through P^-1 = W^T W and P^-1 S P^-1 = W^T G2 W (G2 = H H^T), so with the
// row pass's WHITENED statistics A2 = sum dvar t t^T (A1 = A2 minus the clamped rows), b = sum dmu t (t = W k) and
// g = dLoss/dKL everything is assembled in whitened space (moderate magnitudes) and mapped back by one congruence:
//   dP  = W^T [ A1 - A2 G2 - G2 A2 - 1/2 (beta b^T + b beta^T) + g/2 (I - beta beta^T - G2) ] W
//   dLq = tril(W^T (2 A2 + g I) H) - g diag(1 / L_q,ii),      dm = W^T (b + g beta)
// (no differentiation through the Cholesky factor is needed), then dP goes through K(Z,Z) to d theta and d zf —
// the gradient through the inducing inputs [Z, m_{l-1}].
//
// All matrices are MP x MP row-major (MP = M rounded up to 32, identity/zero padded) and L2-resident; products run
// on the DMMA pipe.  The layers of a model are independent here and are batched in the grid.
#include "common.cuh"

namespace mobo {

constexpr int MAX_BATCH = 8;

struct LayerBatch {
  int n;
  int d, M, MP;
  int kind[MAX_BATCH];
  const double* Zx[MAX_BATCH];
  const double* zf[MAX_BATCH];
  const double* theta[MAX_BATCH];
  const double* m[MAX_BATCH];
  const double* Lq[MAX_BATCH];
  double* ops[MAX_BATCH];
};

struct PtrBatch3 {
  const double* A[MAX_BATCH];
  const double* B[MAX_BATCH];
  double* C[MAX_BATCH];
  double* Ct[MAX_BATCH];   // optional transposed copy of C
};

// ---------------------------------------------------------------------------------------------------
// covariance function between inducing inputs
// ---------------------------------------------------------------------------------------------------
__global__ void kzz_kernel(LayerBatch b, double jitter, int block_index /* which ops block receives P */,
                           int reset_counter /* zero finalize_kernel's arrival counter: ops is caller memory */) {
  __shared__ KernParams kp;
  const int bi = blockIdx.y;
  const int kind = b.kind[bi], d = b.d, M = b.M, MP = b.MP;
  if (reset_counter && blockIdx.x == 0 && threadIdx.x == 0)
    *reinterpret_cast<unsigned long long*>(b.ops[bi] + ops_scal(MP) + SC_COUNTER) = 0ull;
  const double* Zx = b.Zx[bi];
  const double* zf = b.zf[bi];
  if (threadIdx.x == 0) load_kern_params(kp, kind, d, b.theta[bi]);
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= MP * MP) return;
  const int i = idx / MP, j = idx - i * MP;
  double val;
  if (i >= M || j >= M) {
    val = (i == j) ? 1.0 : 0.0;
  } else if (i == j) {
    val = kern_diag(kp, kind == 1 ? zf[i] : 0.0) + jitter;
  } else {
    double D1 = 0.0, D2 = 0.0;
    for (int c = 0; c < d; ++c) {
      const double df = Zx[(size_t)i * d + c] - Zx[(size_t)j * d + c];
      D1 = fma(df * df, kp.il1[c], D1);
      D2 = fma(df * df, kp.il2[c], D2);
    }
    const double E1 = exp(-0.5 * D1);
    if (kind == 0) {
      val = kp.a1 * E1;
    } else {
      const double fi = zf[i], fj = zf[j], dff = fi - fj;
      val = kp.a1 * E1 * (kp.vlin * fi * fj + kp.af * exp(-0.5 * dff * dff * kp.ilf)) + kp.a2 * exp(-0.5 * D2);
    }
  }
  (b.ops[bi] + ops_block(MP, block_index))[idx] = val;
}

// backward through K(Z,Z): dP symmetric.  One warp per inducing point i.
//   dtheta += sum_ij dP_ij dk_ij/dtheta ;  dzf_i = 2 sum_j dP_ij dk(z_i,z_j)/d f_i
constexpr int KZB_WARPS = 8;
constexpr int KZB_NTH = 5 + 2 * kMaxD;
struct KzzBwdBatch {
  const double* dP[MAX_BATCH];
  double* part_theta[MAX_BATCH];   // [gridDim.x][KZB_NTH]
  double* dzf[MAX_BATCH];
};
__global__ void __launch_bounds__(KZB_WARPS * 32) kzz_bwd_kernel(LayerBatch b, KzzBwdBatch o) {
  __shared__ KernParams kp;
  __shared__ double accw[KZB_WARPS][KZB_NTH];
  const int bi = blockIdx.y;
  const int kind = b.kind[bi], d = b.d, M = b.M, MP = b.MP;
  const double* Zx = b.Zx[bi];
  const double* zf = b.zf[bi];
  const double* dP = o.dP[bi];
  if (threadIdx.x == 0) load_kern_params(kp, kind, d, b.theta[bi]);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * KZB_WARPS + warp;
  double th_s[5], th_l1[kMaxD], th_l2[kMaxD];
#pragma unroll
  for (int q = 0; q < 5; ++q) th_s[q] = 0.0;
#pragma unroll
  for (int c = 0; c < kMaxD; ++c) { th_l1[c] = 0.0; th_l2[c] = 0.0; }
  double dzi = 0.0;
  if (i < M) {
    const double fi = kind == 1 ? zf[i] : 0.0;
    double zi[kMaxD];
#pragma unroll
    for (int c = 0; c < kMaxD; ++c) zi[c] = c < d ? Zx[(size_t)i * d + c] : 0.0;
    for (int j = lane; j < M; j += 32) {
      const double gk = dP[(size_t)i * MP + j];
      if (i == j) {
        if (kind == 0) {
          th_s[0] += gk;
        } else {
          th_s[0] += gk * (kp.vlin * fi * fi + kp.af);
          th_s[1] += gk * kp.a1 * fi * fi;
          th_s[2] += gk * kp.a1;
          th_s[4] += gk;
          dzi += gk * kp.a1 * kp.vlin * fi;     // doubled below -> 2 a1 v f
        }
        continue;
      }
      double D1 = 0.0, D2 = 0.0, diff[kMaxD];
#pragma unroll
      for (int c = 0; c < kMaxD; ++c) {
        diff[c] = c < d ? zi[c] - Zx[(size_t)j * d + c] : 0.0;
        D1 = fma(diff[c] * diff[c], kp.il1[c], D1);
        D2 = fma(diff[c] * diff[c], kp.il2[c], D2);
      }
      const double E1 = exp(-0.5 * D1);
      if (kind == 0) {
        const double gkk = gk * kp.a1 * E1;
        th_s[0] = fma(gk, E1, th_s[0]);
#pragma unroll
        for (int c = 0; c < kMaxD; ++c) th_l1[c] = fma(gkk, diff[c] * diff[c], th_l1[c]);
      } else {
        const double fj = zf[j], dff = fi - fj;
        const double Ef = exp(-0.5 * dff * dff * kp.ilf), E2 = exp(-0.5 * D2);
        const double gg = kp.vlin * fi * fj + kp.af * Ef;
        const double s1 = kp.a1 * E1, s2 = kp.a2 * E2;
        const double dEf = s1 * kp.af * Ef * dff * kp.ilf;
        dzi = fma(gk, s1 * kp.vlin * fj - dEf, dzi);
        th_s[0] = fma(gk, E1 * gg, th_s[0]);
        th_s[1] = fma(gk, s1 * fi * fj, th_s[1]);
        th_s[2] = fma(gk, s1 * Ef, th_s[2]);
        th_s[3] = fma(gk, dEf * dff, th_s[3]);
        th_s[4] = fma(gk, E2, th_s[4]);
        const double g1 = gk * s1 * gg, g2 = gk * s2;
#pragma unroll
        for (int c = 0; c < kMaxD; ++c) {
          const double d2 = diff[c] * diff[c];
          th_l1[c] = fma(g1, d2, th_l1[c]);
          th_l2[c] = fma(g2, d2, th_l2[c]);
        }
      }
    }
  }
  dzi = warp_sum(dzi);
  if (lane == 0 && i < M && kind == 1 && o.dzf[bi]) o.dzf[bi][i] = 2.0 * dzi;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const double s = warp_sum(th_s[q]);
    if (lane == 0) accw[warp][q] = s;
  }
#pragma unroll
  for (int c = 0; c < kMaxD; ++c) {
    double s = warp_sum(th_l1[c]);
    if (lane == 0) accw[warp][5 + c] = s;
    s = warp_sum(th_l2[c]);
    if (lane == 0) accw[warp][5 + kMaxD + c] = s;
  }
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid < KZB_NTH) {
    double s = 0.0;
    for (int w = 0; w < KZB_WARPS; ++w) s += accw[w][tid];
    double out = 0.0;
    int slot = -1;
    if (kind == 0) {
      if (tid == 0) { slot = 0; out = s; }
      else if (tid >= 5 && tid < 5 + d) { slot = 1 + (tid - 5); out = s * kp.il1[tid - 5] * sqrt(kp.il1[tid - 5]); }
    } else {
      if (tid < 5) { slot = tid; out = tid == 3 ? s * sqrt(kp.ilf) : s; }
      else if (tid < 5 + d) { slot = tid; out = s * kp.il1[tid - 5] * sqrt(kp.il1[tid - 5]); }
      else if (tid >= 5 + kMaxD && tid < 5 + kMaxD + d) {
        const int c = tid - 5 - kMaxD;
        slot = 5 + d + c;
        out = s * kp.il2[c] * sqrt(kp.il2[c]);
      }
    }
    if (slot >= 0) o.part_theta[bi][(size_t)blockIdx.x * KZB_NTH + slot] = out;
  }
}

// out[bi][i] = sum_b part[bi][b][i]  (fixed order)
struct ReduceBatch { const double* part[MAX_BATCH]; double* out[MAX_BATCH]; int n[MAX_BATCH]; };
__global__ void reduce_batch_kernel(ReduceBatch r, int nblocks, int stride) {
  const int bi = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r.n[bi]) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += r.part[bi][(size_t)b * stride + i];
  r.out[bi][i] = s;
}

// ---------------------------------------------------------------------------------------------------
// 32 x 32 Cholesky factor in registers of one warp (lane = row, columns exchanged by shuffles); used by the
// diagonal CTAs of the operator-chain kernel (opchain.cu).
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ void chol32_inwarp(double (&row)[32], double& rinv, int lane, int& fail) {
  // row[c] = A[lane][c] (lower part valid).  On exit row[c] = L[lane][c] for c <= lane and rinv = 1 / L[lane][lane].
  // Right-looking; the pivot's reciprocal square root (1 ulp) replaces the square root and the division of the
  // textbook form, which sit on the critical path of the whole operator chain.
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double djj = __shfl_sync(0xffffffffu, row[j], j);
    if (!(djj > 0.0)) fail = 1;
    const double r = rsqrt(djj);
    if (lane == j) { row[j] = djj * r; rinv = r; }
    else if (lane > j) row[j] = row[j] * r;
#pragma unroll
    for (int c = j + 1; c < 32; ++c) {
      const double lcj = __shfl_sync(0xffffffffu, row[j], c);   // L[c][j]
      if (lane >= c) row[c] = fma(-row[j], lcj, row[c]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// batched strided GEMM  C = alpha op(A) op(B) + beta C  (n x n x n, n multiple of 32), 32x32 tiles, BK = 32,
// register-prefetched double buffering.  Triangular operands shorten the k range:
//   a_tri: 1 = op(A)[m][k] == 0 for k > m (lower), 2 = == 0 for k < m (upper);  b_tri likewise on op(B)[k][n]:
//   1 = lower (== 0 for n > k), 2 = upper (== 0 for n < k).
// element (m,k) of op(A) = A[m*rsA + k*csA];  element (k,n) of op(B) = B[k*rsB + n*csB]
// ---------------------------------------------------------------------------------------------------
constexpr int GM_T = 32, GM_K = 32, GM_THREADS = 128, GM_LD = GM_T + 4;

struct GemmArgs {
  int n;
  int rsA, csA, rsB, csB;
  int a_tri, b_tri;
  double alpha, beta;
  const double* skip_if_zero[MAX_BATCH];   // optional per-entry device flag: skip the entry when *flag == 0
  PtrBatch3 p;
};

__global__ void __launch_bounds__(GM_THREADS) gemm_kernel(const __grid_constant__ GemmArgs a) {
  if (a.skip_if_zero[blockIdx.z] && *a.skip_if_zero[blockIdx.z] == 0.0) return;
  __shared__ double As[2][GM_K][GM_LD];   // [k][m]
  __shared__ double Bs[2][GM_K][GM_LD];   // [k][n]
  const int bi = blockIdx.z;
  const double* __restrict__ A = a.p.A[bi];
  const double* __restrict__ B = a.p.B[bi];
  double* __restrict__ C = a.p.C[bi];
  double* __restrict__ Ct = a.p.Ct[bi];
  const int n = a.n;
  const int m0 = blockIdx.y * GM_T, n0 = blockIdx.x * GM_T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wi = warp >> 1, wj = warp & 1;          // 16 x 16 per warp
  int kbeg = 0, kend = n;
  if (a.a_tri == 1) kend = min(kend, m0 + GM_T);
  if (a.a_tri == 2) kbeg = max(kbeg, m0);
  if (a.b_tri == 1) kbeg = max(kbeg, n0);
  if (a.b_tri == 2) kend = min(kend, n0 + GM_T);
  double acc[2][2][2] = {};
  // each thread moves 8 elements of each operand per k-tile; index order chosen so global reads coalesce
  double ra[8], rb[8];
  auto gload = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = tid + q * GM_THREADS;
      int kk, mm, kb, nn;
      if (a.csA == 1) { mm = idx >> 5; kk = idx & 31; } else { kk = idx >> 5; mm = idx & 31; }
      if (a.rsB == 1) { nn = idx >> 5; kb = idx & 31; } else { kb = idx >> 5; nn = idx & 31; }
      ra[q] = A[(size_t)(m0 + mm) * a.rsA + (size_t)(k0 + kk) * a.csA];
      rb[q] = B[(size_t)(k0 + kb) * a.rsB + (size_t)(n0 + nn) * a.csB];
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = tid + q * GM_THREADS;
      int kk, mm, kb, nn;
      if (a.csA == 1) { mm = idx >> 5; kk = idx & 31; } else { kk = idx >> 5; mm = idx & 31; }
      if (a.rsB == 1) { nn = idx >> 5; kb = idx & 31; } else { kb = idx >> 5; nn = idx & 31; }
      As[buf][kk][mm] = ra[q];
      Bs[buf][kb][nn] = rb[q];
    }
  };
  if (kbeg < kend) {
    gload(kbeg);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += GM_K) {
      const bool more = k0 + GM_K < kend;
      if (more) gload(k0 + GM_K);
#pragma unroll
      for (int kk = 0; kk < GM_K; kk += 4) {
        double af[2], bf[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) af[x] = As[buf][kk + t][16 * wi + 8 * x + g];
#pragma unroll
        for (int y = 0; y < 2; ++y) bf[y] = Bs[buf][kk + t][16 * wj + 8 * y + g];
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
      }
      if (more) sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 2; ++y)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = m0 + 16 * wi + 8 * x + g, j = n0 + 16 * wj + 8 * y + 2 * t + e;
        const size_t o = (size_t)i * n + j;
        const double v = a.alpha * acc[x][y][e] + (a.beta != 0.0 ? a.beta * C[o] : 0.0);
        C[o] = v;
        if (Ct) Ct[(size_t)j * n + i] = v;
      }
}

struct GemmOperand { const double* p[MAX_BATCH]; bool trans; int tri; };

static int gemm_batched(int nbatch, int n, const GemmOperand& A, const GemmOperand& B, double* const* C,
                        double* const* Ct, double alpha, double beta, const double* const* skip_if_zero,
                        cudaStream_t st) {
  GemmArgs a;
  a.n = n;
  a.rsA = A.trans ? 1 : n; a.csA = A.trans ? n : 1;
  a.rsB = B.trans ? 1 : n; a.csB = B.trans ? n : 1;
  a.a_tri = A.tri; a.b_tri = B.tri;
  a.alpha = alpha; a.beta = beta;
  for (int i = 0; i < MAX_BATCH; ++i) {
    a.p.A[i] = i < nbatch ? A.p[i] : nullptr;
    a.p.B[i] = i < nbatch ? B.p[i] : nullptr;
    a.p.C[i] = i < nbatch ? C[i] : nullptr;
    a.p.Ct[i] = (i < nbatch && Ct) ? Ct[i] : nullptr;
    a.skip_if_zero[i] = (i < nbatch && skip_if_zero) ? skip_if_zero[i] : nullptr;
  }
  dim3 grid(n / GM_T, n / GM_T, nbatch);
  MOBO_LAUNCH("gemm_kernel", st, gemm_kernel<<<grid, GM_THREADS, 0, st>>>(a));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// dm = W^T (b + g beta).  Warp per row of W^T.
struct VecBatch {
  const double* WT[MAX_BATCH]; const double* b[MAX_BATCH]; const double* beta[MAX_BATCH]; const double* dkl[MAX_BATCH];
  double* dm[MAX_BATCH];
};
__global__ void __launch_bounds__(256) white_vec_kernel(VecBatch v, int M, int MP) {
  const int bi = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 8 + warp;
  if (j >= M) return;
  const double* row = v.WT[bi] + (size_t)j * MP;
  const double g = v.dkl[bi][0];
  double s = 0.0;
  for (int i = j + lane; i < MP; i += 32) s = fma(row[i], v.b[bi][i] + g * v.beta[bi][i], s);
  s = warp_sum(s);
  if (lane == 0) v.dm[bi][j] = s;
}

// whitened gradient core  N = A1 - V - V^T - 1/2 (beta b^T + b beta^T) + g/2 (I - beta beta^T - G2)  (dP = W^T N W)
// and F = 2 A2 + g I  (dLq = tril(W^T F H) - g diag(1/Lq_ii));  A1 = A2 - Ac when some row was clamped.
struct CombineBatch {
  const double* A2[MAX_BATCH]; const double* Ac[MAX_BATCH]; const double* V[MAX_BATCH]; const double* G2[MAX_BATCH];
  const double* b[MAX_BATCH]; const double* beta[MAX_BATCH]; const double* dkl[MAX_BATCH];
  const double* clamp_flag[MAX_BATCH]; double* N[MAX_BATCH]; double* F[MAX_BATCH];
};
__global__ void combine_kernel(CombineBatch c, int MP) {
  const int bi = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= MP * MP) return;
  const int i = idx / MP, j = idx - i * MP;
  const size_t tr = (size_t)j * MP + i;
  const double g = c.dkl[bi][0];
  const bool clamped = c.clamp_flag[bi] && *c.clamp_flag[bi] != 0.0;
  const double a2 = 0.5 * (c.A2[bi][idx] + c.A2[bi][tr]);
  const double a1 = clamped ? a2 - 0.5 * (c.Ac[bi][idx] + c.Ac[bi][tr]) : a2;
  const double bi_ = c.b[bi][i], bj_ = c.b[bi][j], be_i = c.beta[bi][i], be_j = c.beta[bi][j];
  const double eye = i == j ? 1.0 : 0.0;
  c.N[bi][idx] = a1 - c.V[bi][idx] - c.V[bi][tr] - 0.5 * (be_i * bj_ + bi_ * be_j) +
                 0.5 * g * (eye - be_i * be_j - c.G2[bi][idx]);
  c.F[bi][idx] = 2.0 * a2 + g * eye;
}

// dLq (M x M, ld M) = tril(E)[:M,:M] - g diag(1 / Lq_ii)
struct DlqBatch { const double* E[MAX_BATCH]; const double* LQ[MAX_BATCH]; const double* dkl[MAX_BATCH]; double* dLq[MAX_BATCH]; };
__global__ void dlq_extract_kernel(DlqBatch q, int M, int MP) {
  const int bi = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * M) return;
  const int i = idx / M, j = idx - i * M;
  double v = j <= i ? q.E[bi][(size_t)i * MP + j] : 0.0;
  if (i == j) v -= q.dkl[bi][0] / q.LQ[bi][(size_t)i * MP + i];
  q.dLq[bi][idx] = v;
}

}  // namespace mobo
