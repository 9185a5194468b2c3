// Fused per-layer row pass of the MFDGP (forward and backward) for sm_100a.
//
// Replaces, for one sparse-GP layer and a batch of R rows, what the reference reaches through
//   MFDGPHiddenLayer.__call__        (mobocmf/layers/mfdgp_hidden_layer.py:245-286)  sample propagation + concat
//   MFDGPHiddenLayer.forward         (mobocmf/layers/mfdgp_hidden_layer.py:232-243)  K(Z_l, X), diag K(X, X)
//   UnwhitenedVariationalStrategy.forward [upstream gpytorch]                        mean / variance of q(f)
// and the autograd backward of those (loss.backward(), mobocmf/util/blackbox_mfdgp_fitter.py:168).
//
// Per row x (with propagated input f) and the per-step operators W = L_p^-1, H = W L_q, beta = W m:
//   k = K(Z_l, x);  t = W k;  u = H^T t;  mu = beta.t;  v = clamp(k_xx - |t|^2, 0) + |u|^2   (train; eval: no clamp)
// which is the reference's  mean = m^T P^-1 k,  var = clamp(k_xx - |L_p^-1 k|^2, 0) + |L_q^T P^-1 k|^2.
// K(Z_l, X) lives only in shared memory (it is written to HBM only when the backward needs it).
//
// Tile: 64 rows x MP inducing points per CTA, 16 warps.  The two triangular M x M products run on the DMMA pipe
// (mma.sync.m8n8k4.f64); each warp owns two 16-row slabs (p, ns-1-p) of the triangular operator so that the
// triangular work is balanced, and streams its slab of the operator from L2 straight into A fragments.
#include "common.cuh"

namespace mobo {

constexpr int TR = 64;             // rows per tile
constexpr int ROW_THREADS = 512;   // 16 warps
constexpr int ROW_WARPS = ROW_THREADS / 32;
constexpr int MAX_MP = 256;        // slab scheme: (MP/32) warp pairs <= 8
constexpr int NCH_MAX = MAX_MP / 32;
constexpr int MAX_THETA = 5 + 2 * kMaxD;

struct RowArgs {
  int kind, d, M, MP;
  const double* Zx;        // M x d inducing inputs (x part)
  const double* zf;        // M: propagated column of the inducing inputs (kind 1): m_{l-1}
  const double* theta;     // constrained kernel hyper-parameters
  const double* ops;       // operator buffer (common.cuh)
  const double* x;         // n x d
  int xrep;                // row r reads x[r / xrep]
  const double* mu_prev;   // previous layer's q(f) moments; row r reads [r / prep]
  const double* var_prev;
  int prep;
  const double* eps;       // normals; row r reads eps[r % eps_mod]
  long long eps_mod;
  const double* f_direct;  // optional explicit propagated input (overrides mu_prev/var_prev/eps)
  long long R;
  int training;            // 1: clamp(k_xx - q, 0) branch; 0: eval branch
  // forward outputs
  double* mu;
  double* var;
  double* craw;            // optional: k_xx - |t|^2 before the clamp (mask for the backward)
  unsigned int* clamp_count; // optional: number of rows whose k_xx - |t|^2 was clamped
  double* Tsave;           // optional row-major [R][MP]: whitened rows t = W k
  double* Usave;           // optional fragment-major [tile][warp][32][32]: u = H^T t
  // backward inputs / outputs
  const double* dmu;
  const double* dvar;
  double* df;              // R: d loss / d f_r   (kind 1)
  double* dxrow;           // optional R x d: d loss / d x of each row
  double* part_theta;      // [grid][MAX_THETA]
  double* part_zf;         // [grid][MP]
  int want_param_grads;
  int want_x_grads;
};

struct RowSmem {
  KernParams kp;
  double red[3][NCH_MAX][TR];     // per warp-pair partial column sums: q1, mu, q2
  double xs[TR][kMaxD];
  double fs[TR];
  double kxx[TR];
  double dmu[TR], dvar[TR], mask[TR];
  double zsT[kMaxD][MAX_MP];
  double zfs[MAX_MP];
  double acc_zf[ROW_WARPS][MAX_MP];   // backward: per-warp d zf accumulators
  double acc_th[ROW_WARPS][MAX_THETA]; // backward: per-warp d theta accumulators (scalars | l1 | l2)
};

__device__ __forceinline__ double* tile_ptr(unsigned char* smem) {
  return reinterpret_cast<double*>(smem + ((sizeof(RowSmem) + 127) / 128) * 128);
}

__host__ inline size_t row_smem_bytes(int MP) {
  return ((sizeof(RowSmem) + 127) / 128) * 128 + (size_t)TR * (MP + 4) * sizeof(double);
}

// One warp: acc(2 slabs x 16 rows x 32 cols) += A[slab rows, k-range] * B[k-range, 32 cols].
// A: row-major MP x MP in global (L2-resident), triangular: LOWER uses k < 16(s+1), UPPER uses k >= 16 s.
// Bs: shared, Bs[col][k] with leading dimension ldb (ldb % 16 == 4 -> conflict-free fragment loads).
template <bool UPPER>
__device__ __forceinline__ void slab_gemm(double (&acc)[2][2][4][2], const double* __restrict__ A, int MP,
                                          const double* Bs, int ldb, int sA, int sB, int half, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int sl = 0; sl < 2; ++sl) {
    const int s = sl == 0 ? sA : sB;
    const int kbeg = UPPER ? 16 * s : 0;
    const int kend = UPPER ? MP : 16 * (s + 1);
    const double* a0p = A + (size_t)(16 * s + g) * MP + t;
    const double* a1p = a0p + (size_t)8 * MP;
    const double* bp = Bs + (size_t)(32 * half + g) * ldb + t;
#pragma unroll 4
    for (int k0 = kbeg; k0 < kend; k0 += 4) {
      const double a0 = __ldg(a0p + k0), a1 = __ldg(a1p + k0);
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) {
        const double b = bp[(size_t)(8 * ct) * ldb + k0];
        dmma884(acc[sl][0][ct][0], acc[sl][0][ct][1], a0, b);
        dmma884(acc[sl][1][ct][0], acc[sl][1][ct][1], a1, b);
      }
    }
  }
}

__device__ __forceinline__ void zero_acc(double (&acc)[2][2][4][2]) {
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) { acc[a][b][c][0] = 0.0; acc[a][b][c][1] = 0.0; }
}

// acc fragment (sl, ib, ct, e) <-> operator row i = 16 s + 8 ib + g, tile column r = 32 half + 8 ct + 2 t + e
__device__ __forceinline__ void store_acc_to_tile(const double (&acc)[2][2][4][2], double* Ks, int ldb, int sA,
                                                  int sB, int half, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int sl = 0; sl < 2; ++sl) {
    const int s = sl == 0 ? sA : sB;
#pragma unroll
    for (int ib = 0; ib < 2; ++ib)
#pragma unroll
      for (int ct = 0; ct < 4; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          Ks[(size_t)(32 * half + 8 * ct + 2 * t + e) * ldb + 16 * s + 8 * ib + g] = acc[sl][ib][ct][e];
  }
}

__device__ __forceinline__ void save_acc_frag(const double (&acc)[2][2][4][2], double* dst, long long tile,
                                              int nact, int wact, int lane) {
  double* p = dst + (((size_t)tile * nact + wact) * 32) * 32 + lane;
#pragma unroll
  for (int sl = 0; sl < 2; ++sl)
#pragma unroll
    for (int ib = 0; ib < 2; ++ib)
#pragma unroll
      for (int ct = 0; ct < 4; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) p[(size_t)(((sl * 2 + ib) * 4 + ct) * 2 + e) * 32] = acc[sl][ib][ct][e];
}

__device__ __forceinline__ void load_acc_frag(double (&acc)[2][2][4][2], const double* src, long long tile,
                                              int nact, int wact, int lane) {
  const double* p = src + (((size_t)tile * nact + wact) * 32) * 32 + lane;
#pragma unroll
  for (int sl = 0; sl < 2; ++sl)
#pragma unroll
    for (int ib = 0; ib < 2; ++ib)
#pragma unroll
      for (int ct = 0; ct < 4; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) acc[sl][ib][ct][e] = p[(size_t)(((sl * 2 + ib) * 4 + ct) * 2 + e) * 32];
}

// loads the rows of one tile: x, propagated input f (sample of the previous layer's q(f), the fused
// reparameterised propagation of layers/mfdgp_hidden_layer.py:263-274), k_xx
__device__ __forceinline__ void load_tile_rows(const RowArgs& a, RowSmem& sm, long long row0, int nvalid) {
  const int tid = threadIdx.x;
  for (int idx = tid; idx < TR * a.d; idx += ROW_THREADS) {
    const int r = idx / a.d, c = idx - r * a.d;
    sm.xs[r][c] = r < nvalid ? a.x[(size_t)((row0 + r) / a.xrep) * a.d + c] : 0.0;
  }
  if (tid < TR) {
    double f = 0.0;
    if (tid < nvalid && a.kind == 1) {
      const long long row = row0 + tid;
      if (a.f_direct) {
        f = a.f_direct[row];
      } else {
        const long long pr = row / a.prep;
        const double vp = fmax(a.var_prev[pr], kMinVariance);
        f = a.mu_prev[pr] + sqrt(vp) * a.eps[row % a.eps_mod];
      }
    }
    sm.fs[tid] = f;
    sm.kxx[tid] = kern_diag(sm.kp, f);
  }
}

__device__ __forceinline__ void load_inducing(const RowArgs& a, RowSmem& sm) {
  const int tid = threadIdx.x;
  if (tid == 0) load_kern_params(sm.kp, a.kind, a.d, a.theta);
  for (int idx = tid; idx < a.MP * a.d; idx += ROW_THREADS) {
    const int j = idx / a.d, c = idx - j * a.d;
    sm.zsT[c][j] = j < a.M ? a.Zx[(size_t)j * a.d + c] : 0.0;
  }
  for (int j = tid; j < a.MP; j += ROW_THREADS) sm.zfs[j] = (a.kind == 1 && j < a.M) ? a.zf[j] : 0.0;
}

__device__ __forceinline__ double kern_pair(const RowSmem& sm, int r, int j) {
  const KernParams& kp = sm.kp;
  double D1 = 0.0, D2 = 0.0;
  for (int c = 0; c < kp.d; ++c) {
    const double df = sm.xs[r][c] - sm.zsT[c][j];
    const double d2 = df * df;
    D1 = fma(d2, kp.il1[c], D1);
    D2 = fma(d2, kp.il2[c], D2);
  }
  const double E1 = exp(-0.5 * D1);
  if (kp.kind == 0) return kp.a1 * E1;
  const double f = sm.fs[r], zf = sm.zfs[j];
  const double dff = f - zf;
  const double Ef = exp(-0.5 * dff * dff * kp.ilf);
  const double E2 = exp(-0.5 * D2);
  return kp.a1 * E1 * (kp.vlin * f * zf + kp.af * Ef) + kp.a2 * E2;
}

__global__ void __launch_bounds__(ROW_THREADS, 1) row_fwd_kernel(const __grid_constant__ RowArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RowSmem& sm = *reinterpret_cast<RowSmem*>(smem_raw);
  double* Ks = tile_ptr(smem_raw);
  const int MP = a.MP, ldb = MP + 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = warp & 1, p = warp >> 1;
  const int npairs = MP / 32, ns = MP / 16;
  const bool active = p < npairs;
  const int sA = p, sB = ns - 1 - p;
  const int nact = 2 * npairs, wact = p * 2 + half;
  const int g = lane >> 2, t = lane & 3;
  const double* W = a.ops + ops_block(MP, OPS_W);
  const double* G = a.ops + ops_block(MP, OPS_HT);
  const double* beta = a.ops + ops_beta(MP);

  load_inducing(a, sm);
  __syncthreads();

  const long long ntiles = (a.R + TR - 1) / TR;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TR;
    const int nvalid = (int)min((long long)TR, a.R - row0);
    load_tile_rows(a, sm, row0, nvalid);
    __syncthreads();
    // ---- K(Z_l, rows) into shared memory ----
    for (int idx = tid; idx < TR * MP; idx += ROW_THREADS) {
      const int r = idx / MP, j = idx - r * MP;
      Ks[(size_t)r * ldb + j] = (j < a.M && r < nvalid) ? kern_pair(sm, r, j) : 0.0;
    }
    __syncthreads();
    // ---- t = W k ----
    double acc[2][2][4][2];
    zero_acc(acc);
    if (active) {
      slab_gemm<false>(acc, W, MP, Ks, ldb, sA, sB, half, lane);
      double pq[4][2], pm[4][2];
#pragma unroll
      for (int ct = 0; ct < 4; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) { pq[ct][e] = 0.0; pm[ct][e] = 0.0; }
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int ib = 0; ib < 2; ++ib) {
          const double bi = __ldg(beta + 16 * (sl == 0 ? sA : sB) + 8 * ib + g);
#pragma unroll
          for (int ct = 0; ct < 4; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const double v = acc[sl][ib][ct][e];
              pq[ct][e] = fma(v, v, pq[ct][e]);
              pm[ct][e] = fma(bi, v, pm[ct][e]);
            }
        }
#pragma unroll
      for (int ct = 0; ct < 4; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double q = pq[ct][e], m = pm[ct][e];
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            q += __shfl_xor_sync(0xffffffffu, q, o);
            m += __shfl_xor_sync(0xffffffffu, m, o);
          }
          if (g == 0) {
            sm.red[0][p][32 * half + 8 * ct + 2 * t + e] = q;
            sm.red[1][p][32 * half + 8 * ct + 2 * t + e] = m;
          }
        }
    }
    __syncthreads();   // every warp is done reading K
    if (active) store_acc_to_tile(acc, Ks, ldb, sA, sB, half, lane);
    __syncthreads();
    if (a.Tsave) {     // whitened rows t = W k, row-major [R][MP], for the backward (SYRK statistics and dt)
      for (int idx = tid; idx < nvalid * MP; idx += ROW_THREADS) {
        const int r = idx / MP, j = idx - r * MP;
        a.Tsave[(size_t)(row0 + r) * MP + j] = Ks[(size_t)r * ldb + j];
      }
    }
    // ---- u = H^T t ----
    if (active) {
      zero_acc(acc);
      slab_gemm<true>(acc, G, MP, Ks, ldb, sA, sB, half, lane);
#pragma unroll
      for (int ct = 0; ct < 4; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double q = 0.0;
#pragma unroll
          for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int ib = 0; ib < 2; ++ib) q = fma(acc[sl][ib][ct][e], acc[sl][ib][ct][e], q);
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
          if (g == 0) sm.red[2][p][32 * half + 8 * ct + 2 * t + e] = q;
        }
      if (a.Usave) save_acc_frag(acc, a.Usave, tile, nact, wact, lane);
    }
    __syncthreads();
    if (tid < nvalid) {
      double q1 = 0.0, mu = 0.0, q2 = 0.0;
      for (int pp = 0; pp < npairs; ++pp) {
        q1 += sm.red[0][pp][tid];
        mu += sm.red[1][pp][tid];
        q2 += sm.red[2][pp][tid];
      }
      const double c = sm.kxx[tid] - q1;
      const double v = (a.training ? fmax(c, 0.0) : c) + q2;
      a.mu[row0 + tid] = mu;
      a.var[row0 + tid] = v;
      if (a.craw) a.craw[row0 + tid] = c;
      if (a.training && c < 0.0 && a.clamp_count) atomicAdd(a.clamp_count, 1u);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// backward of the row pass: given d loss/d mu_r, d loss/d var_r
//   dt = dmu beta - 2 dvar (mask t - H u);  dk = W^T dt
//   then through the covariance function: d theta, d zf (inducing propagated column), d f_r, d x_r.
// The whitened second-order statistics  A2 = sum_r dvar_r t t^T  and  b = sum_r dmu_r t  (t = W k)
// are accumulated by syrk_kernel from the saved T and consumed by the operator backward (matrix_ops.cu).
// ---------------------------------------------------------------------------------------------------
template <bool PARAM, bool XGRAD>
__global__ void __launch_bounds__(ROW_THREADS, 1) row_bwd_kernel(const __grid_constant__ RowArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RowSmem& sm = *reinterpret_cast<RowSmem*>(smem_raw);
  double* Ks = tile_ptr(smem_raw);
  const int MP = a.MP, ldb = MP + 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = warp & 1, p = warp >> 1;
  const int npairs = MP / 32, ns = MP / 16;
  const bool active = p < npairs;
  const int sA = p, sB = ns - 1 - p;
  const int nact = 2 * npairs, wact = p * 2 + half;
  const int g = lane >> 2, t = lane & 3;
  const double* WT = a.ops + ops_block(MP, OPS_WT);
  const double* H = a.ops + ops_block(MP, OPS_H);
  const double* beta = a.ops + ops_beta(MP);
  const int nch = MP / 32;
  const int d = a.d;

  load_inducing(a, sm);
  for (int j = lane; j < MAX_MP; j += 32) sm.acc_zf[warp][j] = 0.0;
  if (lane < MAX_THETA) sm.acc_th[warp][lane] = 0.0;
  __syncthreads();
  const KernParams& kp = sm.kp;

  const long long ntiles = (a.R + TR - 1) / TR;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TR;
    const int nvalid = (int)min((long long)TR, a.R - row0);
    load_tile_rows(a, sm, row0, nvalid);
    if (tid < TR) {
      const bool ok = tid < nvalid;
      sm.dmu[tid] = ok ? a.dmu[row0 + tid] : 0.0;
      sm.dvar[tid] = ok ? a.dvar[row0 + tid] : 0.0;
      sm.mask[tid] = (ok && a.training && a.craw) ? (a.craw[row0 + tid] >= 0.0 ? 1.0 : 0.0) : 1.0;
    }
    double acc[2][2][4][2];
    // ---- u tile -> shared (B layout) ----
    if (active) {
      load_acc_frag(acc, a.Usave, tile, nact, wact, lane);
      store_acc_to_tile(acc, Ks, ldb, sA, sB, half, lane);
    }
    __syncthreads();
    // ---- y = H u ;  dt = dmu beta - 2 dvar (mask t - y) ----
    if (active) {
      zero_acc(acc);
      slab_gemm<false>(acc, H, MP, Ks, ldb, sA, sB, half, lane);
    }
    __syncthreads();   // u is no longer needed: stage the saved t tile through shared memory
    for (int idx = tid; idx < TR * MP; idx += ROW_THREADS) {
      const int r = idx / MP, j = idx - r * MP;
      Ks[(size_t)r * ldb + j] = r < nvalid ? a.Tsave[(size_t)(row0 + r) * MP + j] : 0.0;
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int ib = 0; ib < 2; ++ib) {
          const int i = 16 * (sl == 0 ? sA : sB) + 8 * ib + g;
          const double bi = __ldg(beta + i);
#pragma unroll
          for (int ct = 0; ct < 4; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int r = 32 * half + 8 * ct + 2 * t + e;
              const double tv = Ks[(size_t)r * ldb + i];
              acc[sl][ib][ct][e] = sm.dmu[r] * bi - 2.0 * sm.dvar[r] * (sm.mask[r] * tv - acc[sl][ib][ct][e]);
            }
        }
    }
    __syncthreads();
    if (active) store_acc_to_tile(acc, Ks, ldb, sA, sB, half, lane);
    __syncthreads();
    // ---- dk = W^T dt ----
    if (active) {
      zero_acc(acc);
      slab_gemm<true>(acc, WT, MP, Ks, ldb, sA, sB, half, lane);
    }
    __syncthreads();
    if (active) store_acc_to_tile(acc, Ks, ldb, sA, sB, half, lane);
    __syncthreads();
    // ---- through the covariance function: warp <-> rows, lanes <-> inducing points of a 32-chunk ----
    double th_s[5], th_l1[kMaxD], th_l2[kMaxD], azf[NCH_MAX];
#pragma unroll
    for (int i = 0; i < 5; ++i) th_s[i] = 0.0;
#pragma unroll
    for (int c = 0; c < kMaxD; ++c) { th_l1[c] = 0.0; th_l2[c] = 0.0; }
#pragma unroll
    for (int ch = 0; ch < NCH_MAX; ++ch) azf[ch] = 0.0;
    for (int r = warp; r < nvalid; r += ROW_WARPS) {
      const double f = sm.fs[r];
      double rdf = 0.0;
      double rdx[kMaxD];
#pragma unroll
      for (int c = 0; c < kMaxD; ++c) rdx[c] = 0.0;
#pragma unroll
      for (int ch = 0; ch < NCH_MAX; ++ch) {
        const int j = 32 * ch + lane;
        if (ch < nch && j < a.M) {
          const double gk = Ks[(size_t)r * ldb + j];
          double D1 = 0.0, D2 = 0.0;
          double diff[kMaxD];
#pragma unroll
          for (int c = 0; c < kMaxD; ++c) {
            diff[c] = c < d ? sm.xs[r][c] - sm.zsT[c][j] : 0.0;
            const double d2 = diff[c] * diff[c];
            D1 = fma(d2, kp.il1[c], D1);
            D2 = fma(d2, kp.il2[c], D2);
          }
          const double E1 = exp(-0.5 * D1);
          if (kp.kind == 0) {
            const double gkk = gk * kp.a1 * E1;
            if (PARAM) {
              th_s[0] = fma(gk, E1, th_s[0]);
#pragma unroll
              for (int c = 0; c < kMaxD; ++c) th_l1[c] = fma(gkk, diff[c] * diff[c], th_l1[c]);
            }
            if (XGRAD) {
#pragma unroll
              for (int c = 0; c < kMaxD; ++c) rdx[c] = fma(-gkk * diff[c], kp.il1[c], rdx[c]);
            }
          } else {
            const double zf = sm.zfs[j];
            const double dff = f - zf;
            const double Ef = exp(-0.5 * dff * dff * kp.ilf);
            const double E2 = exp(-0.5 * D2);
            const double gg = kp.vlin * f * zf + kp.af * Ef;   // k_lin + k_f
            const double s1 = kp.a1 * E1, s2 = kp.a2 * E2;
            const double dEf = s1 * kp.af * Ef * dff * kp.ilf;  // s1 af Ef (f - zf) / lf^2
            rdf = fma(gk, s1 * kp.vlin * zf - dEf, rdf);
            const double g1 = gk * s1 * gg, g2 = gk * s2;
            if (PARAM) {
              azf[ch] = fma(gk, s1 * kp.vlin * f + dEf, azf[ch]);
              th_s[0] = fma(gk, E1 * gg, th_s[0]);          // a1
              th_s[1] = fma(gk, s1 * f * zf, th_s[1]);      // v
              th_s[2] = fma(gk, s1 * Ef, th_s[2]);          // af
              th_s[3] = fma(gk, dEf * dff, th_s[3]);        // lf (x 1/lf applied at flush)
              th_s[4] = fma(gk, E2, th_s[4]);               // a2
#pragma unroll
              for (int c = 0; c < kMaxD; ++c) {
                const double d2 = diff[c] * diff[c];
                th_l1[c] = fma(g1, d2, th_l1[c]);           // (x 1/l^3 applied at flush)
                th_l2[c] = fma(g2, d2, th_l2[c]);
              }
            }
            if (XGRAD) {
#pragma unroll
              for (int c = 0; c < kMaxD; ++c) rdx[c] -= diff[c] * (g1 * kp.il1[c] + g2 * kp.il2[c]);
            }
          }
        }
      }
      // row-wise sums over the inducing points
      const double dvm = sm.dvar[r] * sm.mask[r];
      if (kp.kind == 1) {
        rdf = warp_sum(rdf);
        if (lane == 0) a.df[row0 + r] = rdf + dvm * 2.0 * kp.a1 * kp.vlin * f;   // + d k_xx / d f
      }
      if (XGRAD) {
#pragma unroll
        for (int c = 0; c < kMaxD; ++c)
          if (c < d) {
            const double s = warp_sum(rdx[c]);
            if (lane == 0) a.dxrow[(size_t)(row0 + r) * d + c] = s;
          }
      }
      // d k_xx / d theta (diag term of the variance), added once per row by lane 0
      if (PARAM && lane == 0) {
        if (kp.kind == 0) {
          th_s[0] += dvm;
        } else {
          th_s[0] += dvm * (kp.vlin * f * f + kp.af);
          th_s[1] += dvm * kp.a1 * f * f;
          th_s[2] += dvm * kp.a1;
          th_s[4] += dvm;
        }
      }
    }
    if (PARAM) {
      // per-warp accumulators in shared memory (each slot has one owner -> deterministic)
#pragma unroll
      for (int ch = 0; ch < NCH_MAX; ++ch)
        if (ch < nch) sm.acc_zf[warp][32 * ch + lane] += azf[ch];
      double red;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        red = warp_sum(th_s[i]);
        if (lane == 0) sm.acc_th[warp][i] += red;
      }
#pragma unroll
      for (int c = 0; c < kMaxD; ++c)
        if (c < d) {
          red = warp_sum(th_l1[c]);
          if (lane == 0) sm.acc_th[warp][5 + c] += red;
          red = warp_sum(th_l2[c]);
          if (lane == 0) sm.acc_th[warp][5 + kMaxD + c] += red;
        }
    }
    __syncthreads();
  }
  if (PARAM) {
    // ---- flush the per-CTA partials (fixed summation order -> deterministic) ----
    __syncthreads();
    for (int j = tid; j < MP; j += ROW_THREADS) {
      double s = 0.0;
      for (int w = 0; w < ROW_WARPS; ++w) s += sm.acc_zf[w][j];
      a.part_zf[(size_t)blockIdx.x * MP + j] = s;
    }
    if (tid < MAX_THETA) {
      double s = 0.0;
      for (int w = 0; w < ROW_WARPS; ++w) s += sm.acc_th[w][tid];
      // map (scalars | l1 | l2) accumulators to the theta layout and apply the 1/l^3 factors
      double out = 0.0;
      int slot = -1;
      if (kp.kind == 0) {
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
      if (slot >= 0) a.part_theta[(size_t)blockIdx.x * MAX_THETA + slot] = out;
    }
  }
}

// sums the per-CTA partials in a fixed order: out[i] = sum_b part[b][i]
__global__ void reduce_partials_kernel(const double* __restrict__ part, int nblocks, int n, int stride,
                                       double* __restrict__ out, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += part[(size_t)b * stride + i];
  out[i] = accumulate ? out[i] + s : s;
}

// ---------------------------------------------------------------------------------------------------
// A2 = sum_r w_r t_r t_r^T (lower 64x64 tiles) from the saved whitened rows T [R][MP]; split over row chunks, partial
// tiles reduced in a fixed order by syrk_reduce_kernel which also mirrors to the full symmetric matrix.
// w_r = dvar_r (which = 0) or dvar_r * [clamped row] (which = 1, only when some row was clamped).
// ---------------------------------------------------------------------------------------------------
constexpr int SY_T = 64, SY_K = 32, SY_THREADS = 256;

__global__ void __launch_bounds__(SY_THREADS) syrk_kernel(const double* __restrict__ K, const double* __restrict__ dvar,
                                                          const double* __restrict__ craw, int which, int MP,
                                                          long long R, int nchunk, double* __restrict__ part,
                                                          const unsigned int* __restrict__ clamp_count,
                                                          const double* __restrict__ dmu,
                                                          double* __restrict__ part_alpha) {
  if (which == 1 && (clamp_count == nullptr || *clamp_count == 0u)) return;
  __shared__ double As[SY_K][SY_T + 4];   // As[r][i] (scaled by w_r)
  __shared__ double Bs[SY_K][SY_T + 4];   // Bs[r][j]
  const int nb = MP / SY_T + (MP % SY_T ? 1 : 0);
  // tile index -> (bi >= bj)
  int tix = blockIdx.x, bi = 0;
  while (tix > bi) { tix -= bi + 1; ++bi; }
  const int bj = tix;
  const int chunk = blockIdx.y;
  const long long rows_per = ((R + nchunk - 1) / nchunk + SY_K - 1) / SY_K * SY_K;
  const long long rbeg = (long long)chunk * rows_per, rend = min(R, rbeg + rows_per);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wi = warp >> 1, wj = warp & 1;        // warp tile: 16 (i) x 32 (j)
  double acc[2][4][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) { acc[x][y][0] = 0.0; acc[x][y][1] = 0.0; }
  const bool do_alpha = which == 0 && bi == bj && part_alpha != nullptr;   // dalpha_j = sum_r dmu_r K[r][j]
  double al = 0.0;
  for (long long r0 = rbeg; r0 < rend; r0 += SY_K) {
    for (int idx = tid; idx < SY_K * SY_T; idx += SY_THREADS) {
      const int r = idx / SY_T, c = idx - r * SY_T;
      const long long row = r0 + r;
      double va = 0.0, vb = 0.0;
      if (row < rend) {
        double w = dvar[row];
        if (which == 1) w = (craw[row] < 0.0) ? w : 0.0;
        const int ci = bi * SY_T + c, cj = bj * SY_T + c;
        if (ci < MP) va = w * K[(size_t)row * MP + ci];
        if (cj < MP) vb = K[(size_t)row * MP + cj];
        if (do_alpha) al = fma(dmu[row], vb, al);
      }
      As[r][c] = va;
      Bs[r][c] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < SY_K; k0 += 4) {
      double af[2], bf[4];
#pragma unroll
      for (int x = 0; x < 2; ++x) af[x] = As[k0 + t][16 * wi + 8 * x + g];
#pragma unroll
      for (int y = 0; y < 4; ++y) bf[y] = Bs[k0 + t][32 * wj + 8 * y + g];
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
    }
    __syncthreads();
  }
  if (do_alpha) {   // tid % 64 is this thread's fixed column; fold the 4 row-phases in a fixed order
    As[tid / SY_T][tid % SY_T] = al;
    __syncthreads();
    if (tid < SY_T && bi * SY_T + tid < MP)
      part_alpha[(size_t)chunk * MP + bi * SY_T + tid] = (As[0][tid] + As[1][tid]) + (As[2][tid] + As[3][tid]);
  }
  double* out = part + ((size_t)chunk * gridDim.x + blockIdx.x) * SY_T * SY_T;
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        out[(size_t)(16 * wi + 8 * x + g) * SY_T + 32 * wj + 8 * y + 2 * t + e] = acc[x][y][e];
  (void)nb;
}

__global__ void syrk_reduce_kernel(const double* __restrict__ part, int ntiles, int nchunk, int MP,
                                   double* __restrict__ A, int which, const unsigned int* __restrict__ clamp_count,
                                   double* __restrict__ clamp_flag_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0 && which == 1 && clamp_flag_out) *clamp_flag_out = clamp_count ? (double)(*clamp_count) : 0.0;
  if (idx >= MP * MP) return;
  if (which == 1 && (clamp_count == nullptr || *clamp_count == 0u)) { A[idx] = 0.0; return; }
  int i = idx / MP, j = idx - (idx / MP) * MP;
  if (j > i) { const int tmp = i; i = j; j = tmp; }
  const int bi = i / SY_T, bj = j / SY_T;
  const int tile = bi * (bi + 1) / 2 + bj;
  const int li = i - bi * SY_T, lj = j - bj * SY_T;
  double s = 0.0;
  for (int c = 0; c < nchunk; ++c) s += part[((size_t)c * ntiles + tile) * SY_T * SY_T + (size_t)li * SY_T + lj];
  A[idx] = s;
}

// ---------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

int row_grid(long long R) {
  const long long ntiles = (R + TR - 1) / TR;
  return (int)(ntiles < (long long)num_sms() ? (ntiles > 0 ? ntiles : 1) : num_sms());
}

int launch_row_fwd(const RowArgs& a, cudaStream_t st) {
  if (a.MP % 32 != 0 || a.MP > MAX_MP || a.d > kMaxD || a.M > a.MP) return -2;
  const size_t smem = row_smem_bytes(a.MP);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(row_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem_bytes(MAX_MP));
    attr_done = true;
  }
  if (a.R <= 0) return 0;
  MOBO_LAUNCH("row_fwd_kernel", st, row_fwd_kernel<<<row_grid(a.R), ROW_THREADS, smem, st>>>(a));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_row_bwd(const RowArgs& a, int grid, cudaStream_t st) {
  if (a.MP % 32 != 0 || a.MP > MAX_MP || a.d > kMaxD || a.M > a.MP) return -2;
  const size_t smem = row_smem_bytes(a.MP);
  static bool attr_done = false;
  if (!attr_done) {
    const int mx = (int)row_smem_bytes(MAX_MP);
    cudaFuncSetAttribute(row_bwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(row_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(row_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    attr_done = true;
  }
  if (a.R <= 0) return 0;
  if (a.want_param_grads && a.want_x_grads) {
    MOBO_LAUNCH("row_bwd_kernel<param,x>", st, row_bwd_kernel<true, true><<<grid, ROW_THREADS, smem, st>>>(a));
  } else if (a.want_x_grads) {
    MOBO_LAUNCH("row_bwd_kernel<x>", st, row_bwd_kernel<false, true><<<grid, ROW_THREADS, smem, st>>>(a));
  } else {
    MOBO_LAUNCH("row_bwd_kernel<param>", st, row_bwd_kernel<true, false><<<grid, ROW_THREADS, smem, st>>>(a));
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_reduce_partials(const double* part, int nblocks, int n, int stride, double* out, int accumulate,
                           cudaStream_t st) {
  MOBO_LAUNCH("reduce_partials_kernel", st, reduce_partials_kernel<<<(n + 127) / 128, 128, 0, st>>>(part, nblocks, n, stride, out, accumulate));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int syrk_ntiles(int MP) {
  const int nb = (MP + SY_T - 1) / SY_T;
  return nb * (nb + 1) / 2;
}

int syrk_nchunk(int MP, long long R) {
  const int nt = syrk_ntiles(MP);
  long long want = (2LL * num_sms() + nt - 1) / nt;
  const long long maxc = (R + 4 * SY_K - 1) / (4 * SY_K);
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  return (int)want;
}

size_t syrk_part_doubles(int MP, long long R) { return (size_t)syrk_ntiles(MP) * syrk_nchunk(MP, R) * SY_T * SY_T; }

// part: syrk_part_doubles(MP, R) doubles of scratch; part_alpha: nchunk * MP doubles (which == 0 only)
int launch_syrk(const double* K, const double* dvar, const double* craw, int which, int MP, long long R,
                double* part, double* A, const unsigned int* clamp_count, const double* dmu, double* part_alpha,
                double* dalpha, double* clamp_flag_out, cudaStream_t st) {
  const int nt = syrk_ntiles(MP), nc = syrk_nchunk(MP, R);
  dim3 grid(nt, nc);
  MOBO_LAUNCH("syrk_kernel", st, syrk_kernel<<<grid, SY_THREADS, 0, st>>>(K, dvar, craw, which, MP, R, nc, part, clamp_count, dmu,
                                           which == 0 ? part_alpha : nullptr));
  MOBO_LAUNCH("syrk_reduce_kernel", st, syrk_reduce_kernel<<<(MP * MP + 255) / 256, 256, 0, st>>>(part, nt, nc, MP, A, which, clamp_count, clamp_flag_out));
  if (which == 0 && part_alpha && dalpha)
    MOBO_LAUNCH("reduce_partials_kernel", st, reduce_partials_kernel<<<(MP + 127) / 128, 128, 0, st>>>(part_alpha, nc, MP, MP, dalpha, 0));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mobo
