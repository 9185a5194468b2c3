// Per-step M x M operator chain of one sparse-GP layer (forward and backward) for sm_100a.
//
// Forward (what the reference reaches through UnwhitenedVariationalStrategy.forward / kl_mvn_mvn [upstream gpytorch],
// called from mobocmf/layers/mfdgp_hidden_layer.py:286 and mobocmf/mlls/variational_elbo_mf.py:40):
//   P = K(Z_l, Z_l) + jitter I      (CovarianceMatrixMF.add_jitter, layers/mfdgp_hidden_layer.py:17-20)
//   L = chol(P);  W = L^-1;  H = W tril(L_q);  beta = W m;  alpha = W^T beta
//   KL = 1/2 [ 2 sum log L_ii - sum log L_q,ii^2 + |beta|^2 + |H|_F^2 - M ]
// Backward: from the row pass's second-order statistics (A2 = sum dvar k k^T, Ac = clamped-row part, dalpha) and
// dKL to d m, d L_q, d theta, d zf (the gradient through the inducing inputs [Z, m_{l-1}]).
//
// All matrices are MP x MP row-major (MP = M rounded up to 32, identity/zero padded) and L2-resident; products run
// on the DMMA pipe through one strided 64x64-tile kernel.
#include "common.cuh"

namespace mobo {

constexpr int MAX_MP_FINAL = 256;

// ---------------------------------------------------------------------------------------------------
// covariance function between inducing inputs
// ---------------------------------------------------------------------------------------------------
__global__ void kzz_kernel(int kind, int d, int M, int MP, const double* __restrict__ Zx,
                           const double* __restrict__ zf, const double* __restrict__ theta, double jitter,
                           double* __restrict__ P) {
  __shared__ KernParams kp;
  if (threadIdx.x == 0) load_kern_params(kp, kind, d, theta);
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
  P[idx] = val;
}

// backward through K(Z,Z): dP symmetric.  One warp per inducing point i.
//   dtheta += sum_ij dP_ij dk_ij/dtheta ;  dzf_i += 2 sum_j dP_ij dk(z_i,z_j)/d f_i
constexpr int KZB_WARPS = 8;
__global__ void __launch_bounds__(KZB_WARPS * 32) kzz_bwd_kernel(int kind, int d, int M, int MP,
                                                                 const double* __restrict__ Zx,
                                                                 const double* __restrict__ zf,
                                                                 const double* __restrict__ theta,
                                                                 const double* __restrict__ dP,
                                                                 double* __restrict__ part_theta,   // [grid][5+2*kMaxD]
                                                                 double* __restrict__ dzf_out, int accumulate_zf) {
  __shared__ KernParams kp;
  __shared__ double accw[KZB_WARPS][5 + 2 * kMaxD];
  if (threadIdx.x == 0) load_kern_params(kp, kind, d, theta);
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
        diff[c] = c < d ? Zx[(size_t)i * d + c] - Zx[(size_t)j * d + c] : 0.0;
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
  if (lane == 0 && i < M && kind == 1) dzf_out[i] = (accumulate_zf ? dzf_out[i] : 0.0) + 2.0 * dzi;
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
  if (tid < 5 + 2 * kMaxD) {
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
    if (slot >= 0) part_theta[(size_t)blockIdx.x * (5 + 2 * kMaxD) + slot] = out;
  }
}

// ---------------------------------------------------------------------------------------------------
// Cholesky  L = chol(P)   (blocked right-looking, one CTA per matrix, operands stay in L2)
// ---------------------------------------------------------------------------------------------------
constexpr int CH_NB = 32, CH_THREADS = 1024, CH_LDP = 36;

__global__ void __launch_bounds__(CH_THREADS, 1) chol_kernel(const double* __restrict__ P, double* __restrict__ L,
                                                            int MP, double* __restrict__ scal) {
  extern __shared__ __align__(16) double sh[];
  double (*D)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(sh);
  double* panel = sh + CH_NB * (CH_NB + 1) + 8;   // [(MP-32)][CH_LDP]
  __shared__ int fail;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (tid == 0) fail = 0;
  for (int idx = tid; idx < MP * MP; idx += CH_THREADS) {
    const int i = idx / MP, j = idx - i * MP;
    L[idx] = j <= i ? P[idx] : 0.0;
  }
  __syncthreads();
  for (int k0 = 0; k0 < MP; k0 += CH_NB) {
    D[tid >> 5][tid & 31] = L[(size_t)(k0 + (tid >> 5)) * MP + k0 + (tid & 31)];
    __syncthreads();
    if (warp == 0) {
      for (int j = 0; j < CH_NB; ++j) {
        const double djj = D[j][j];
        if (!(djj > 0.0)) { if (lane == 0) fail = 1; }
        const double ljj = sqrt(djj);
        if (lane == j) D[j][j] = ljj;
        if (lane > j) D[lane][j] = D[lane][j] / ljj;
        __syncwarp();
        if (lane > j) {
          const double lij = D[lane][j];
          for (int c = j + 1; c <= lane; ++c) D[lane][c] -= lij * D[c][j];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    {
      const int i = tid >> 5, j = tid & 31;
      L[(size_t)(k0 + i) * MP + k0 + j] = j <= i ? D[i][j] : 0.0;
    }
    const int nrows = MP - k0 - CH_NB;
    if (nrows > 0) {
      // panel: X D^T = A  -> forward substitution along each row
      if (tid < nrows) {
        double x[CH_NB];
        const double* arow = L + (size_t)(k0 + CH_NB + tid) * MP + k0;
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) x[c] = arow[c];
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) {
          double s = x[c];
#pragma unroll
          for (int k = 0; k < c; ++k) s = fma(-x[k], D[c][k], s);
          x[c] = s / D[c][c];
        }
        double* orow = L + (size_t)(k0 + CH_NB + tid) * MP + k0;
#pragma unroll
        for (int c = 0; c < CH_NB; ++c) {
          orow[c] = x[c];
          panel[(size_t)tid * CH_LDP + c] = x[c];
        }
      }
      __syncthreads();
      // trailing update: L[i][j] -= panel[i] . panel[j]  on 32x32 blocks (bi >= bj), DMMA
      const int nblk = nrows / CH_NB, nb2 = nblk * (nblk + 1) / 2;
      for (int blk = warp; blk < nb2; blk += CH_THREADS / 32) {
        int bi = 0, rem = blk;
        while (rem > bi) { rem -= bi + 1; ++bi; }
        const int bj = rem;
        double acc[4][4][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) { acc[x][y][0] = 0.0; acc[x][y][1] = 0.0; }
#pragma unroll
        for (int kk = 0; kk < CH_NB; kk += 4) {
          double af[4], bf[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) af[x] = panel[(size_t)(bi * CH_NB + 8 * x + g) * CH_LDP + kk + t];
#pragma unroll
          for (int y = 0; y < 4; ++y) bf[y] = panel[(size_t)(bj * CH_NB + 8 * y + g) * CH_LDP + kk + t];
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
        }
        double* base = L + (size_t)(k0 + CH_NB + bi * CH_NB) * MP + k0 + CH_NB + bj * CH_NB;
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int ii = 8 * x + g, jj = 8 * y + 2 * t + e;
              if (bi > bj || jj <= ii) base[(size_t)ii * MP + jj] -= acc[x][y][e];
            }
      }
    }
    __syncthreads();
  }
  if (tid == 0) scal[SC_STATUS] = fail ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// W = L^-1  (blocked, one CTA per matrix): diagonal 32x32 blocks by substitution, then block diagonals
// W_ik = -W_ii (sum_{j=k}^{i-1} L_ij W_jk) in order of increasing distance i-k, DMMA products.
// ---------------------------------------------------------------------------------------------------
constexpr int TI_THREADS = 1024;
__global__ void __launch_bounds__(TI_THREADS, 1) trtri_kernel(const double* __restrict__ L, double* __restrict__ W,
                                                             int MP) {
  extern __shared__ __align__(16) double sh[];
  const int nb = MP / 32;
  double* Sblk = sh;                       // [nb][32][36]  S = sum_j L_ij W_jk  per block of the current diagonal
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int idx = tid; idx < MP * MP; idx += TI_THREADS) W[idx] = 0.0;
  __syncthreads();
  // diagonal blocks: lane = column c of the block, solve L_bb x = e_c
  if (warp < nb) {
    const int b = warp, c = lane;
    const double* Lb = L + (size_t)(32 * b) * MP + 32 * b;
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; ++k) s = fma(-__ldg(Lb + (size_t)i * MP + k), x[k], s);
      x[i] = (i >= c) ? s / __ldg(Lb + (size_t)i * MP + i) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) W[(size_t)(32 * b + i) * MP + 32 * b + c] = x[i];
  }
  __threadfence_block();
  __syncthreads();
  for (int dd = 1; dd < nb; ++dd) {
    const int nblocks = nb - dd;
    // 4 warps per block, each a 16x16 quadrant of the 32x32 block
    // phase A: S = sum_j L_ij W_jk
    for (int item = warp; item < nblocks * 4; item += TI_THREADS / 32) {
      const int bq = item >> 2, q = item & 3;
      const int k = bq, i = bq + dd;
      const int r0 = 16 * (q >> 1), c0 = 16 * (q & 1);
      double acc[2][2][2] = {};
      for (int j = k; j < i; ++j) {
        const double* Lij = L + (size_t)(32 * i) * MP + 32 * j;
        const double* Wjk = W + (size_t)(32 * j) * MP + 32 * k;
#pragma unroll
        for (int kk = 0; kk < 32; kk += 4) {
          double af[2], bf[2];
#pragma unroll
          for (int x = 0; x < 2; ++x) af[x] = Lij[(size_t)(r0 + 8 * x + g) * MP + kk + t];
#pragma unroll
          for (int y = 0; y < 2; ++y) bf[y] = Wjk[(size_t)(kk + t) * MP + c0 + 8 * y + g];
#pragma unroll
          for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int y = 0; y < 2; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
        }
      }
      double* S = Sblk + (size_t)bq * 32 * 36;
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 2; ++y)
#pragma unroll
          for (int e = 0; e < 2; ++e) S[(size_t)(r0 + 8 * x + g) * 36 + c0 + 8 * y + 2 * t + e] = acc[x][y][e];
    }
    __syncthreads();
    // phase B: W_ik = -W_ii S
    for (int item = warp; item < nblocks * 4; item += TI_THREADS / 32) {
      const int bq = item >> 2, q = item & 3;
      const int k = bq, i = bq + dd;
      const int r0 = 16 * (q >> 1), c0 = 16 * (q & 1);
      const double* Wii = W + (size_t)(32 * i) * MP + 32 * i;
      const double* S = Sblk + (size_t)bq * 32 * 36;
      double acc[2][2][2] = {};
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
        double af[2], bf[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) af[x] = Wii[(size_t)(r0 + 8 * x + g) * MP + kk + t];
#pragma unroll
        for (int y = 0; y < 2; ++y) bf[y] = S[(size_t)(kk + t) * 36 + c0 + 8 * y + g];
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
      }
      double* Wik = W + (size_t)(32 * i) * MP + 32 * k;
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 2; ++y)
#pragma unroll
          for (int e = 0; e < 2; ++e) Wik[(size_t)(r0 + 8 * x + g) * MP + c0 + 8 * y + 2 * t + e] = -acc[x][y][e];
    }
    __threadfence_block();
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// strided small GEMM:  C = alpha * op(A) op(B) + beta * C   (n x n x n, n = MP multiple of 32), 64x64 tiles
// element (m,k) of op(A) = A[m*rsA + k*csA];  element (k,n) of op(B) = B[k*rsB + n*csB]
// ---------------------------------------------------------------------------------------------------
constexpr int GM_T = 64, GM_K = 16, GM_THREADS = 256;
__global__ void __launch_bounds__(GM_THREADS) gemm_kernel(int n, const double* __restrict__ A, int rsA, int csA,
                                                         const double* __restrict__ B, int rsB, int csB,
                                                         double* __restrict__ C, double alpha, double beta) {
  __shared__ double As[GM_K][GM_T + 4];
  __shared__ double Bs[GM_K][GM_T + 4];
  const int m0 = blockIdx.y * GM_T, n0 = blockIdx.x * GM_T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wi = warp >> 1, wj = warp & 1;
  double acc[2][4][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) { acc[x][y][0] = 0.0; acc[x][y][1] = 0.0; }
  for (int k0 = 0; k0 < n; k0 += GM_K) {
    for (int idx = tid; idx < GM_K * GM_T; idx += GM_THREADS) {
      int kk, mm;
      if (csA == 1) { mm = idx / GM_K; kk = idx - mm * GM_K; } else { kk = idx / GM_T; mm = idx - kk * GM_T; }
      As[kk][mm] = (m0 + mm < n) ? A[(size_t)(m0 + mm) * rsA + (size_t)(k0 + kk) * csA] : 0.0;
      int kb, nn;
      if (rsB == 1) { nn = idx / GM_K; kb = idx - nn * GM_K; } else { kb = idx / GM_T; nn = idx - kb * GM_T; }
      Bs[kb][nn] = (n0 + nn < n) ? B[(size_t)(k0 + kb) * rsB + (size_t)(n0 + nn) * csB] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GM_K; kk += 4) {
      double af[2], bf[4];
#pragma unroll
      for (int x = 0; x < 2; ++x) af[x] = As[kk + t][16 * wi + 8 * x + g];
#pragma unroll
      for (int y = 0; y < 4; ++y) bf[y] = Bs[kk + t][32 * wj + 8 * y + g];
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = m0 + 16 * wi + 8 * x + g, j = n0 + 32 * wj + 8 * y + 2 * t + e;
        if (i < n && j < n) {
          const size_t o = (size_t)i * n + j;
          C[o] = alpha * acc[x][y][e] + (beta != 0.0 ? beta * C[o] : 0.0);
        }
      }
}

static int gemm(int n, const double* A, bool tA, const double* B, bool tB, double* C, double alpha, double beta,
                cudaStream_t st) {
  dim3 grid((n + GM_T - 1) / GM_T, (n + GM_T - 1) / GM_T);
  MOBO_LAUNCH("gemm_kernel", st, gemm_kernel<<<grid, GM_THREADS, 0, st>>>(n, A, tA ? 1 : n, tA ? n : 1, B, tB ? 1 : n, tB ? n : 1, C, alpha, beta));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------
// elementwise helpers on MP x MP blocks
// ---------------------------------------------------------------------------------------------------
enum EwOp {
  EW_TRANSPOSE = 0,      // out = in^T
  EW_PAD_TRIL = 1,       // out(MP) = tril(in(M x M, ld M)) zero padded
  EW_SUB = 2,            // out = in - in2
  EW_SCALE = 3,          // out = s * in         (s read from device scalar * hs)
  EW_TRIL_INPLACE = 4,   // out = tril(in)
  EW_NEG_TRIL_DIAG = 5,  // out = -tril(in) + diag(s / L_ii)   (in2 = L)
  EW_PHI = 6,            // out = tril(in) with halved diagonal
  EW_SYM = 7,            // out = 1/2 (in + in^T)
  EW_RANK1_ADD = 8,      // out += u v^T  (u = vec1, v = vec2)
};
__global__ void ew_kernel(int op, int M, int MP, const double* __restrict__ in, const double* __restrict__ in2,
                          double* __restrict__ out, const double* __restrict__ sdev, double hs,
                          const double* __restrict__ v1, const double* __restrict__ v2) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= MP * MP) return;
  const int i = idx / MP, j = idx - i * MP;
  const double s = sdev ? sdev[0] * hs : hs;
  switch (op) {
    case EW_TRANSPOSE: out[idx] = in[(size_t)j * MP + i]; break;
    case EW_PAD_TRIL: out[idx] = (i < M && j <= i) ? in[(size_t)i * M + j] : 0.0; break;
    case EW_SUB: out[idx] = in[idx] - in2[idx]; break;
    case EW_SCALE: out[idx] = s * in[idx]; break;
    case EW_TRIL_INPLACE: out[idx] = j <= i ? in[idx] : 0.0; break;
    case EW_NEG_TRIL_DIAG:
      out[idx] = (j <= i ? -in[idx] : 0.0) + ((i == j && i < M) ? s / in2[idx] : 0.0);
      break;
    case EW_PHI: out[idx] = j < i ? in[idx] : (j == i ? 0.5 * in[idx] : 0.0); break;
    case EW_SYM: out[idx] = 0.5 * (in[idx] + in[(size_t)j * MP + i]); break;
    case EW_RANK1_ADD: out[idx] += v1[i] * v2[j]; break;
  }
}

static int ew(int op, int M, int MP, const double* in, const double* in2, double* out, const double* sdev, double hs,
              const double* v1, const double* v2, cudaStream_t st) {
  MOBO_LAUNCH("ew_kernel", st, ew_kernel<<<(MP * MP + 255) / 256, 256, 0, st>>>(op, M, MP, in, in2, out, sdev, hs, v1, v2));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// beta = W m, alpha = W^T beta, KL pieces.  One CTA, warp-per-row GEMVs (coalesced over the row).
__global__ void __launch_bounds__(256) finalize_kernel(int M, int MP, const double* __restrict__ m, double* ops) {
  __shared__ double sb[MAX_MP_FINAL];
  __shared__ double red[4][8];
  const double* L = ops + ops_block(MP, OPS_L);
  const double* W = ops + ops_block(MP, OPS_W);
  const double* WT = ops + ops_block(MP, OPS_WT);
  const double* H = ops + ops_block(MP, OPS_H);
  const double* LQ = ops + ops_block(MP, OPS_LQ);
  double* beta = ops + ops_beta(MP);
  double* alpha = ops + ops_alpha(MP);
  double* scal = ops + ops_scal(MP);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = warp; i < MP; i += 8) {
    double s = 0.0;
    if (i < M)
      for (int k = lane; k <= i; k += 32) s = fma(W[(size_t)i * MP + k], m[k], s);
    s = warp_sum(s);
    if (lane == 0) { sb[i] = s; beta[i] = s; }
  }
  __syncthreads();
  for (int j = warp; j < MP; j += 8) {
    double s = 0.0;
    for (int i = j + lane; i < MP; i += 32) s = fma(WT[(size_t)j * MP + i], sb[i], s);
    s = warp_sum(s);
    if (lane == 0) alpha[j] = s;
  }
  double b2 = 0.0, ldp = 0.0, ldq = 0.0, h2 = 0.0;
  for (int j = tid; j < MP; j += 256) {
    b2 = fma(sb[j], sb[j], b2);
    if (j < M) {
      ldp += 2.0 * log(L[(size_t)j * MP + j]);
      const double q = LQ[(size_t)j * MP + j];
      ldq += log(q * q);
    }
  }
  for (int idx = tid; idx < MP * MP; idx += 256) h2 = fma(H[idx], H[idx], h2);
  b2 = warp_sum(b2); ldp = warp_sum(ldp); ldq = warp_sum(ldq); h2 = warp_sum(h2);
  if (lane == 0) { red[0][warp] = b2; red[1][warp] = ldp; red[2][warp] = ldq; red[3][warp] = h2; }
  __syncthreads();
  if (tid == 0) {
    double v[4];
    for (int q = 0; q < 4; ++q) { v[q] = 0.0; for (int w = 0; w < 8; ++w) v[q] += red[q][w]; }
    scal[SC_BETA2] = v[0]; scal[SC_LOGDET_P] = v[1]; scal[SC_LOGDET_Q] = v[2]; scal[SC_H2] = v[3];
    scal[SC_KL] = 0.5 * (v[1] - v[2] + v[0] + v[3] - (double)M);
  }
}

// dbeta = W dalpha + dkl beta ; dm = W^T dbeta.  One CTA.  Writes dbeta (MP) and dm (M).
__global__ void __launch_bounds__(256) dbeta_kernel(int M, int MP, const double* __restrict__ ops,
                                                   const double* __restrict__ dalpha, const double* __restrict__ dkl,
                                                   double* __restrict__ dbeta_out, double* __restrict__ dm) {
  __shared__ double sb[MAX_MP_FINAL];
  const double* W = ops + ops_block(MP, OPS_W);
  const double* WT = ops + ops_block(MP, OPS_WT);
  const double* beta = ops + ops_beta(MP);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double gk = dkl[0];
  for (int i = warp; i < MP; i += 8) {
    double s = 0.0;
    for (int k = lane; k <= i; k += 32) s = fma(W[(size_t)i * MP + k], dalpha[k], s);
    s = warp_sum(s);
    if (lane == 0) { s += gk * beta[i]; sb[i] = s; dbeta_out[i] = s; }
  }
  __syncthreads();
  for (int j = warp; j < M; j += 8) {
    double s = 0.0;
    for (int i = j + lane; i < MP; i += 32) s = fma(WT[(size_t)j * MP + i], sb[i], s);
    s = warp_sum(s);
    if (lane == 0) dm[j] = s;
  }
}

// dLq (M x M, ld M) = tril(X)[:M,:M] - dkl * diag(1 / Lq_ii)
__global__ void dlq_extract_kernel(int M, int MP, const double* __restrict__ X, const double* __restrict__ LQ,
                                   const double* __restrict__ dkl, double* __restrict__ dLq) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * M) return;
  const int i = idx / M, j = idx - i * M;
  double v = j <= i ? X[(size_t)i * MP + j] : 0.0;
  if (i == j) v -= dkl[0] / LQ[(size_t)i * MP + i];
  dLq[idx] = v;
}

}  // namespace mobo
