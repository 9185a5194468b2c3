// extern "C" entry points of libmobocmf_b200.so (declared in include/mobocmf_b200.h) and the host-side kernel
// sequences behind them.
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "../../include/mobocmf_b200.h"
#include "matrix_ops.cu"
#include "opchain.cu"
#include "row_pass.cu"
#include "step.cu"
#include "rff.cu"

using namespace mobo;

static inline int padded(int M) { return ((M + 31) / 32) * 32; }
#define MOBO_TRY(x) do { int _e = (x); if (_e) return _e; } while (0)

extern "C" {

int mobo_abi_version(void) { return 102; }

long long mobo_launch_count(void) { return prof_state().launches; }

void mobo_profile_enable(int on) { prof_state().on = on != 0; }

int mobo_profile_collect(char* names, size_t names_bytes, float* ms, int max_records) {
  ProfState& s = prof_state();
  int n = 0;
  size_t off = 0;
  for (ProfRec& r : s.recs) {
    cudaEventSynchronize(r.e1);
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    if (n < max_records) {
      const size_t len = strlen(r.name) + 1;
      if (off + len <= names_bytes) {
        memcpy(names + off, r.name, len);
        off += len;
        ms[n++] = t;
      }
    }
    s.pool.push_back(r.e0);
    s.pool.push_back(r.e1);
  }
  s.recs.clear();
  return n;
}
int mobo_padded_m(int M) { return padded(M); }
size_t mobo_ops_doubles(int M) { return ops_size(padded(M)); }

size_t mobo_rows_save_doubles(int M, long long R) {
  const long long ntiles = (R + TR - 1) / TR;
  return (size_t)ntiles * TR * padded(M);
}

size_t mobo_rows_bwd_work_doubles(int M, long long R) {
  const int MP = padded(M);
  return (size_t)256 * KG_CTAS_PER_SM * (MAX_THETA + MP) + syrk_part_doubles(MP, R) + syrk_alpha_doubles(MP, syrk_nchunk(MP, R)) + 64 +
         mobo_rows_save_doubles(M, R);   // + the dk scratch [R][MP]
}

size_t mobo_precompute_bwd_work_doubles(int M) {
  const int MP = padded(M);
  return (size_t)8 * MP * MP + (size_t)((M + KZB_WARPS - 1) / KZB_WARPS) * KZB_NTH + 64;
}

static int fill_batch(LayerBatch& b, int nl, const int* kinds, int d, int M, const double* const* Zx,
                      const double* const* zf, const double* const* theta, const double* const* m,
                      const double* const* Lq, double* const* ops) {
  if (nl < 1 || nl > MAX_BATCH) return -2;
  b.n = nl; b.d = d; b.M = M; b.MP = padded(M);
  if (d > kMaxD || b.MP > MAX_MP) return -2;
  for (int i = 0; i < MAX_BATCH; ++i) {
    const bool ok = i < nl;
    b.kind[i] = ok ? kinds[i] : 0;
    b.Zx[i] = ok ? Zx[i] : nullptr;
    b.zf[i] = (ok && zf) ? zf[i] : nullptr;
    b.theta[i] = ok ? theta[i] : nullptr;
    b.m[i] = (ok && m) ? m[i] : nullptr;
    b.Lq[i] = (ok && Lq) ? Lq[i] : nullptr;
    b.ops[i] = (ok && ops) ? ops[i] : nullptr;
  }
  return 0;
}

int mobo_kzz(int kind, int d, int M, const double* Zx, const double* zf, const double* theta, double jitter,
             double* P, void* stream) {
  // P is written where block OPS_P of an operator buffer starting at (P - OPS_P * MP^2) would live
  LayerBatch b;
  const int MP = padded(M);
  double* fake_ops = P - ops_block(MP, OPS_P);
  MOBO_TRY(fill_batch(b, 1, &kind, d, M, &Zx, &zf, &theta, nullptr, nullptr, &fake_ops));
  cudaStream_t st = (cudaStream_t)stream;
  MOBO_LAUNCH("kzz_kernel", st, kzz_kernel<<<dim3((MP * MP + 255) / 256, 1), 256, 0, st>>>(b, jitter, OPS_P, 0));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

static int model_precompute(int nl, const int* kinds, int d, int M, const double* const* Zx, const double* const* zf,
                            const double* const* theta, const double* const* m, const double* const* Lq,
                            double jitter, double* const* ops, void* stream, bool reset_flags);

int mobo_model_precompute(int nl, const int* kinds, int d, int M, const double* const* Zx, const double* const* zf,
                          const double* const* theta, const double* const* m, const double* const* Lq,
                          double jitter, double* const* ops, void* stream) {
  return model_precompute(nl, kinds, d, M, Zx, zf, theta, m, Lq, jitter, ops, stream, true);
}

// reset_flags = false: the caller has already zeroed the flags region of every operator buffer on this stream
static int model_precompute(int nl, const int* kinds, int d, int M, const double* const* Zx, const double* const* zf,
                            const double* const* theta, const double* const* m, const double* const* Lq,
                            double jitter, double* const* ops, void* stream, bool reset_flags) {
  cudaStream_t st = (cudaStream_t)stream;
  LayerBatch b;
  MOBO_TRY(fill_batch(b, nl, kinds, d, M, Zx, zf, theta, m, Lq, ops));
  const int MP = b.MP;
  // one cooperative launch: a CTA per 32 x 32 block of the lower triangle of every layer (opchain.cu)
  const int nb = MP / 32, nblk = nb * (nb + 1) / 2;
  if (reset_flags) MOBO_LAUNCH("opchain_reset_kernel", st, opchain_reset_kernel<<<nl, 128, 0, st>>>(b));
  void* args[] = {(void*)&b, (void*)&jitter};
  prof_begin("opchain_kernel", st);
  const cudaError_t e = cudaLaunchCooperativeKernel((const void*)opchain_kernel, dim3(nl * nblk), dim3(OC_THREADS), args,
                                                    0, st);
  prof_end(st);
  if (e != cudaSuccess) return -1;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mobo_model_precompute_bwd(int nl, const int* kinds, int d, int M, const double* const* Zx,
                              const double* const* zf, const double* const* theta, const double* const* m,
                              const double* const* Lq, const double* const* ops, const double* const* gops,
                              double* const* work, double* const* dtheta, double* const* dzf, double* const* dm,
                              double* const* dLq, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LayerBatch b;
  MOBO_TRY(fill_batch(b, nl, kinds, d, M, Zx, zf, theta, m, Lq, const_cast<double* const*>(ops)));
  const int MP = b.MP;
  const size_t MP2 = (size_t)MP * MP;
  enum { T_G2 = 0, T_V, T_NN, T_F, T_X1, T_DP, T_X2, T_Y2, T_N };
  auto blk = [&](int i, int which) { return work[i] + (size_t)which * MP2; };
  auto mk = [&](GemmOperand& o, bool trans, int tri, auto getter) {
    o.trans = trans; o.tri = tri;
    for (int i = 0; i < MAX_BATCH; ++i) o.p[i] = i < nl ? getter(i) : nullptr;
  };
  double* C[MAX_BATCH];
  auto outs = [&](int which) { for (int i = 0; i < MAX_BATCH; ++i) C[i] = i < nl ? blk(i, which) : nullptr; };
  GemmOperand A, B;
  // G2 = H H^T
  mk(A, false, 1, [&](int i) { return ops[i] + ops_block(MP, OPS_H); });
  mk(B, true, 2, [&](int i) { return ops[i] + ops_block(MP, OPS_H); });
  outs(T_G2);
  MOBO_TRY(gemm_batched(nl, MP, A, B, C, nullptr, 1.0, 0.0, nullptr, st));
  // V = A2 G2
  mk(A, false, 0, [&](int i) { return gops[i] + ops_block(MP, OPS_W); });
  mk(B, false, 0, [&](int i) { return (const double*)blk(i, T_G2); });
  outs(T_V);
  MOBO_TRY(gemm_batched(nl, MP, A, B, C, nullptr, 1.0, 0.0, nullptr, st));
  // whitened core N and F = 2 A2 + g I
  {
    CombineBatch c;
    for (int i = 0; i < MAX_BATCH; ++i) {
      const bool ok = i < nl;
      c.A2[i] = ok ? gops[i] + ops_block(MP, OPS_W) : nullptr;
      c.Ac[i] = ok ? gops[i] + ops_block(MP, OPS_H) : nullptr;
      c.V[i] = ok ? blk(i, T_V) : nullptr; c.G2[i] = ok ? blk(i, T_G2) : nullptr;
      c.b[i] = ok ? gops[i] + ops_alpha(MP) : nullptr;
      c.beta[i] = ok ? ops[i] + ops_beta(MP) : nullptr;
      c.dkl[i] = ok ? gops[i] + ops_scal(MP) + SC_KL : nullptr;
      c.clamp_flag[i] = ok ? gops[i] + ops_scal(MP) + SC_CLAMP : nullptr;
      c.N[i] = ok ? blk(i, T_NN) : nullptr; c.F[i] = ok ? blk(i, T_F) : nullptr;
    }
    MOBO_LAUNCH("combine_kernel", st, combine_kernel<<<dim3((MP2 + 255) / 256, nl), 256, 0, st>>>(c, MP));
  }
  // dP = W^T N W
  mk(A, false, 2, [&](int i) { return ops[i] + ops_block(MP, OPS_WT); });
  mk(B, false, 0, [&](int i) { return (const double*)blk(i, T_NN); });
  outs(T_X1);
  MOBO_TRY(gemm_batched(nl, MP, A, B, C, nullptr, 1.0, 0.0, nullptr, st));
  mk(A, false, 0, [&](int i) { return (const double*)blk(i, T_X1); });
  mk(B, false, 1, [&](int i) { return ops[i] + ops_block(MP, OPS_W); });
  outs(T_DP);
  MOBO_TRY(gemm_batched(nl, MP, A, B, C, nullptr, 1.0, 0.0, nullptr, st));
  // dLq = tril(W^T F H) - g diag(1 / Lq_ii)
  mk(A, false, 0, [&](int i) { return (const double*)blk(i, T_F); });
  mk(B, false, 1, [&](int i) { return ops[i] + ops_block(MP, OPS_H); });
  outs(T_X2);
  MOBO_TRY(gemm_batched(nl, MP, A, B, C, nullptr, 1.0, 0.0, nullptr, st));
  mk(A, false, 2, [&](int i) { return ops[i] + ops_block(MP, OPS_WT); });
  mk(B, false, 0, [&](int i) { return (const double*)blk(i, T_X2); });
  outs(T_Y2);
  MOBO_TRY(gemm_batched(nl, MP, A, B, C, nullptr, 1.0, 0.0, nullptr, st));
  {
    DlqBatch q;
    for (int i = 0; i < MAX_BATCH; ++i) {
      const bool ok = i < nl;
      q.E[i] = ok ? blk(i, T_Y2) : nullptr; q.LQ[i] = ok ? ops[i] + ops_block(MP, OPS_LQ) : nullptr;
      q.dkl[i] = ok ? gops[i] + ops_scal(MP) + SC_KL : nullptr; q.dLq[i] = ok ? dLq[i] : nullptr;
    }
    MOBO_LAUNCH("dlq_extract_kernel", st,
                dlq_extract_kernel<<<dim3((M * M + 255) / 256, nl), 256, 0, st>>>(q, M, MP));
  }
  // dm = W^T (b + g beta)
  {
    VecBatch v;
    for (int i = 0; i < MAX_BATCH; ++i) {
      const bool ok = i < nl;
      v.WT[i] = ok ? ops[i] + ops_block(MP, OPS_WT) : nullptr;
      v.b[i] = ok ? gops[i] + ops_alpha(MP) : nullptr;
      v.beta[i] = ok ? ops[i] + ops_beta(MP) : nullptr;
      v.dkl[i] = ok ? gops[i] + ops_scal(MP) + SC_KL : nullptr;
      v.dm[i] = ok ? dm[i] : nullptr;
    }
    MOBO_LAUNCH("white_vec_kernel", st, white_vec_kernel<<<dim3((M + 7) / 8, nl), 256, 0, st>>>(v, M, MP));
  }
  // through K(Z, Z)
  {
    const int nblk = (M + KZB_WARPS - 1) / KZB_WARPS;
    KzzBwdBatch o;
    ReduceBatch r;
    for (int i = 0; i < MAX_BATCH; ++i) {
      const bool ok = i < nl;
      o.dP[i] = ok ? blk(i, T_DP) : nullptr;
      o.part_theta[i] = ok ? work[i] + T_N * MP2 : nullptr;
      o.dzf[i] = (ok && dzf) ? dzf[i] : nullptr;
      r.part[i] = o.part_theta[i];
      r.out[i] = ok ? dtheta[i] : nullptr;
      r.n[i] = ok ? theta_size(kinds[i], d) : 0;
    }
    MOBO_LAUNCH("kzz_bwd_kernel", st, kzz_bwd_kernel<<<dim3(nblk, nl), KZB_WARPS * 32, 0, st>>>(b, o));
    MOBO_LAUNCH("reduce_batch_kernel", st, reduce_batch_kernel<<<dim3(1, nl), 32, 0, st>>>(r, nblk, KZB_NTH));
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mobo_layer_precompute(int kind, int d, int M, const double* Zx, const double* zf, const double* theta,
                          const double* m, const double* Lq, double jitter, double* ops, void* stream) {
  return mobo_model_precompute(1, &kind, d, M, &Zx, &zf, &theta, &m, &Lq, jitter, &ops, stream);
}

int mobo_layer_precompute_bwd(int kind, int d, int M, const double* Zx, const double* zf, const double* theta,
                              const double* m, const double* Lq, const double* ops, const double* gops,
                              double* work, double* dtheta, double* dzf, double* dm, double* dLq, void* stream) {
  return mobo_model_precompute_bwd(1, &kind, d, M, &Zx, &zf, &theta, &m, &Lq, &ops, &gops, &work, &dtheta, &dzf, &dm,
                                   &dLq, stream);
}

static void fill_row_args(RowArgs& a, int kind, int d, int M, const double* Zx, const double* zf,
                          const double* theta, const double* ops, const double* x, int xrep, const double* mu_prev,
                          const double* var_prev, int prep, const double* eps, long long eps_mod,
                          const double* f_direct, long long R, int training) {
  a.kind = kind; a.d = d; a.M = M; a.MP = padded(M);
  a.Zx = Zx; a.zf = zf; a.theta = theta; a.ops = ops; a.x = x; a.xrep = xrep < 1 ? 1 : xrep;
  a.mu_prev = mu_prev; a.var_prev = var_prev; a.prep = prep < 1 ? 1 : prep;
  a.eps = eps; a.eps_mod = eps_mod < 1 ? 1 : eps_mod; a.f_direct = f_direct; a.R = R; a.training = training;
  a.mu = nullptr; a.var = nullptr; a.craw = nullptr; a.clamp_count = nullptr;
  a.Tsave = nullptr; a.Usave = nullptr;
  a.dmu = nullptr; a.dvar = nullptr; a.dk = nullptr; a.df = nullptr; a.dxrow = nullptr; a.part_theta = nullptr; a.part_zf = nullptr;
  a.want_param_grads = 0; a.want_x_grads = 0;
}

int mobo_layer_rows_fwd(int kind, int d, int M, const double* Zx, const double* zf, const double* theta,
                        const double* ops, const double* x, int xrep, const double* mu_prev, const double* var_prev,
                        int prep, const double* eps, long long eps_mod, const double* f_direct, long long R,
                        int training, double* mu, double* var, double* craw, unsigned int* clamp_count,
                        double* Tsave, double* Usave, void* stream) {
  RowArgs a;
  fill_row_args(a, kind, d, M, Zx, zf, theta, ops, x, xrep, mu_prev, var_prev, prep, eps, eps_mod, f_direct, R,
                training);
  a.mu = mu; a.var = var; a.craw = craw; a.clamp_count = clamp_count;
  a.Tsave = Tsave; a.Usave = Usave;
  return launch_row_fwd(a, (cudaStream_t)stream);
}

// the whitened statistics of one layer's backward (SYRK): A2, Ac, b into the operator-gradient buffer
static int rows_bwd_stats(int MP, long long R, int training, const double* Tsave, const double* dmu,
                          const double* dvar, const double* craw, const unsigned int* clamp_count, double* part_syrk,
                          double* part_alpha, double* gops, cudaStream_t st) {
  MOBO_TRY(launch_syrk(Tsave, dvar, craw, 0, MP, R, part_syrk, gops + ops_block(MP, OPS_W), clamp_count, dmu,
                       part_alpha, gops + ops_alpha(MP), nullptr, st));
  MOBO_TRY(launch_syrk(Tsave, dvar, craw, 1, MP, R, part_syrk, gops + ops_block(MP, OPS_H),
                       training ? clamp_count : nullptr, nullptr, nullptr, nullptr,
                       gops + ops_scal(MP) + SC_CLAMP, st));
  return 0;
}

// product kernel + covariance-gradient kernel + the fixed-order reduction of the latter's per-CTA partials
// the fold of the covariance-gradient kernel's per-CTA partials (latency-bound: a few hundred dependent loads per
// output) may run on another stream `rs` once `ev` (recorded on `st` behind the kernel) has fired
static int rows_bwd_main(RowArgs& a, int kind, int d, int M, long long R, double* dk, double* part, double* dtheta,
                         double* dzf, cudaStream_t st, cudaStream_t rs = nullptr, cudaEvent_t ev = nullptr) {
  const int MP = a.MP;
  const int grid = kgrad_grid(R);
  a.dk = dk;
  a.part_theta = part;
  a.part_zf = part + (size_t)grid * MAX_THETA;
  MOBO_TRY(launch_row_bwd(a, st));
  if (a.want_param_grads && dtheta) {
    cudaStream_t r = st;
    if (rs && ev) {
      if (cudaEventRecord(ev, st) != cudaSuccess || cudaStreamWaitEvent(rs, ev, 0) != cudaSuccess) return -1;
      r = rs;
    }
    MOBO_TRY(launch_reduce_partials(a.part_theta, grid, theta_size(kind, d), MAX_THETA, dtheta, 0, r));
    if (kind == 1 && dzf) MOBO_TRY(launch_reduce_partials(a.part_zf, grid, M, MP, dzf, 0, r));
  }
  return 0;
}

static void fill_row_bwd_args(RowArgs& a, const double* dmu, const double* dvar, const double* craw,
                              const double* Tsave, const double* Usave, int want_param_grads, double* df,
                              double* dxrow) {
  a.dmu = dmu; a.dvar = dvar; a.craw = const_cast<double*>(craw);
  a.Tsave = const_cast<double*>(Tsave); a.Usave = const_cast<double*>(Usave);
  a.df = df; a.dxrow = dxrow;
  a.want_param_grads = want_param_grads; a.want_x_grads = dxrow != nullptr;
  if (!a.want_param_grads && !a.want_x_grads) a.want_param_grads = 1;
  a.sm_reserve = 0;
}

int mobo_layer_rows_bwd(int kind, int d, int M, const double* Zx, const double* zf, const double* theta,
                        const double* ops, const double* x, int xrep, const double* mu_prev, const double* var_prev,
                        int prep, const double* eps, long long eps_mod, const double* f_direct, long long R,
                        int training, const double* dmu, const double* dvar, const double* craw,
                        const unsigned int* clamp_count, const double* Tsave,
                        const double* Usave, int want_param_grads, double* df, double* dxrow, double* dtheta,
                        double* dzf, double* gops, double* work, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RowArgs a;
  fill_row_args(a, kind, d, M, Zx, zf, theta, ops, x, xrep, mu_prev, var_prev, prep, eps, eps_mod, f_direct, R,
                training);
  fill_row_bwd_args(a, dmu, dvar, craw, Tsave, Usave, want_param_grads, df, dxrow);
  const int MP = a.MP;
  const int grid = kgrad_grid(R);
  double* dk = work;                                             // [R][MP] scratch, first so that it stays aligned
  double* part = work + mobo_rows_save_doubles(M, R);
  double* part_syrk = part + (size_t)grid * (MAX_THETA + MP);
  double* part_alpha = part_syrk + syrk_part_doubles(MP, R);
  MOBO_TRY(rows_bwd_main(a, kind, d, M, R, dk, part, dtheta, dzf, st));
  if (gops) MOBO_TRY(rows_bwd_stats(MP, R, training, Tsave, dmu, dvar, craw, clamp_count, part_syrk, part_alpha, gops, st));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------
// fused ELBO step
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct StepLayout {
  size_t theta[ST_MAX_LAYERS], ops[ST_MAX_LAYERS], gops[ST_MAX_LAYERS], mu[ST_MAX_LAYERS], var[ST_MAX_LAYERS],
      craw[ST_MAX_LAYERS], dmu[ST_MAX_LAYERS], dvar[ST_MAX_LAYERS], df[ST_MAX_LAYERS], Ts[ST_MAX_LAYERS],
      Us[ST_MAX_LAYERS], dtheta_rows[ST_MAX_LAYERS], dtheta_pre[ST_MAX_LAYERS], dzf_rows[ST_MAX_LAYERS],
      dzf_pre[ST_MAX_LAYERS], dm_pre[ST_MAX_LAYERS], dLq_tmp[ST_MAX_LAYERS], pre_work[ST_MAX_LAYERS],
      ell_part[ST_MAX_LAYERS], stats0[ST_MAX_LAYERS], stats1[ST_MAX_LAYERS], stats_alpha[ST_MAX_LAYERS],
      kg_part[ST_MAX_LAYERS];
  long long R[ST_MAX_LAYERS];
  int ell_blocks[ST_MAX_LAYERS];
  size_t noise, clamp, rows_work, total;
};

StepLayout step_layout(int L, int d, int M, int S, long long B) {
  StepLayout y;
  size_t off = 0;
  auto take = [&](size_t n) { const size_t o = off; off += (n + 15) / 16 * 16; return o; };
  const int MP = padded(M);
  y.noise = take(ST_MAX_LAYERS);
  y.clamp = take(ST_MAX_LAYERS);
  size_t rows_work = 0;
  for (int l = 0; l < L; ++l) {
    const long long R = l == 0 ? B : B * (long long)S;
    y.R[l] = R;
    y.ell_blocks[l] = (int)((R + ELL_THREADS - 1) / ELL_THREADS);
    y.theta[l] = take(ST_MAX_THETA);
    y.ops[l] = take(ops_size(MP));
    y.gops[l] = take(ops_size(MP));
    y.mu[l] = take(R); y.var[l] = take(R); y.craw[l] = take(R);
    y.dmu[l] = take(R); y.dvar[l] = take(R); y.df[l] = take(R);
    y.Ts[l] = take(mobo_rows_save_doubles(M, R));
    y.Us[l] = take(mobo_rows_save_doubles(M, R));
    y.dtheta_rows[l] = take(ST_MAX_THETA); y.dtheta_pre[l] = take(ST_MAX_THETA);
    y.dzf_rows[l] = take(MP); y.dzf_pre[l] = take(MP); y.dm_pre[l] = take(MP);
    y.dLq_tmp[l] = take((size_t)M * M);
    y.pre_work[l] = take(mobo_precompute_bwd_work_doubles(M));
    y.ell_part[l] = take(2 * (size_t)y.ell_blocks[l]);
    const size_t w = mobo_rows_save_doubles(M, R) + (size_t)256 * KG_CTAS_PER_SM * (MAX_THETA + MP) + 64;
    rows_work = w > rows_work ? w : rows_work;
    // SYRK partials per layer and per weight set: their fold runs on the side stream while the main stream is already
    // in the next layer's SYRK
    y.stats0[l] = take(syrk_part_doubles(MP, R));
    y.stats1[l] = take(syrk_part_doubles(MP, R));
    y.stats_alpha[l] = take(syrk_alpha_doubles(MP, syrk_nchunk(MP, R)) + 64);
    // per-CTA partials of the covariance-gradient kernel, per layer: their fold runs on the side stream too
    y.kg_part[l] = take((size_t)256 * KG_CTAS_PER_SM * (MAX_THETA + MP) + 64);
  }
  y.rows_work = take(rows_work);
  y.total = off;
  (void)d;
  return y;
}

}  // namespace

// Side stream of the fused step: the fold of the SYRK partials and the operator-chain backward of layer l run there
// while the main stream already works on layer l - 1's row kernels; forked and joined with events inside
// mobo_elbo_step, so the pattern is also valid under CUDA-graph capture.  The stream and its events belong to a step
// context the caller creates (mobo_step_ctx_create: one per concurrently trained model, so that independent models do
// not serialise on each other's side stream); a step without a context uses one per-device default, created under a
// lock on first use - the only state this library keeps.
namespace {
struct SideCtx {
  int device = -1; cudaStream_t s = nullptr; cudaEvent_t fork[ST_MAX_LAYERS]; cudaEvent_t join = nullptr;
  cudaEvent_t done[ST_MAX_LAYERS];      // layer l's operator-chain backward (its d L_q, d m, d theta shares) is complete
  cudaEvent_t kg[ST_MAX_LAYERS];        // layer l's covariance-gradient kernel is complete (its partials may be folded)
};
bool side_ctx_init(SideCtx& c) {
  if (cudaGetDevice(&c.device) != cudaSuccess) return false;
  if (cudaStreamCreateWithFlags(&c.s, cudaStreamNonBlocking) != cudaSuccess) return false;
  for (int i = 0; i < ST_MAX_LAYERS; ++i)
    if (cudaEventCreateWithFlags(&c.fork[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.done[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.kg[i], cudaEventDisableTiming) != cudaSuccess) return false;
  return cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) == cudaSuccess;
}
SideCtx* default_side_ctx() {
  static SideCtx ctx[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  SideCtx& c = ctx[dev & 63];
  if (c.device < 0 && !side_ctx_init(c)) return nullptr;
  return &c;
}
}  // namespace

void* mobo_step_ctx_create(void) {
  SideCtx* c = new SideCtx();
  if (!side_ctx_init(*c)) { delete c; return nullptr; }
  return c;
}
int mobo_step_ctx_wait_layer(void* ctx, int layer, void* stream) {
  SideCtx* c = static_cast<SideCtx*>(ctx);
  if (!c || layer < 0 || layer >= ST_MAX_LAYERS) return -2;
  return cudaStreamWaitEvent((cudaStream_t)stream, c->done[layer], 0) == cudaSuccess ? 0 : -1;
}
void mobo_step_ctx_destroy(void* ctx) {
  SideCtx* c = static_cast<SideCtx*>(ctx);
  if (!c) return;
  if (c->s) cudaStreamDestroy(c->s);
  for (int i = 0; i < ST_MAX_LAYERS; ++i) { if (c->fork[i]) cudaEventDestroy(c->fork[i]); if (c->done[i]) cudaEventDestroy(c->done[i]); if (c->kg[i]) cudaEventDestroy(c->kg[i]); }
  if (c->join) cudaEventDestroy(c->join);
  delete c;
}

static bool g_side_stream_on = true;
// SMs the backward product kernel (one persistent CTA per SM, all of its registers and shared memory) leaves to the
// side stream's operator-chain backward; measured on C4 (ms / step): round 1: 0 -> 2.84, 4 -> 2.72, 8 -> 2.71, 16 -> 2.75;
// round-2 kernels: 0 -> 2.51, 4 -> 2.41, 8 -> 2.44, 12 -> 2.45, 16 -> 2.45
constexpr int kSideStreamSMs = 4;
static int side_stream_sms() {      // development override: MOBO_SIDE_SMS=n
  static const int n = getenv("MOBO_SIDE_SMS") ? atoi(getenv("MOBO_SIDE_SMS")) : kSideStreamSMs;
  return n;
}
void mobo_step_side_stream(int on) { g_side_stream_on = on != 0; }

size_t mobo_elbo_step_workspace_doubles(int L, int d, int M, int S, long long B) {
  if (L < 1 || L > ST_MAX_LAYERS) return 0;
  return step_layout(L, d, M, S, B).total;
}

int mobo_elbo_step(const mobo_step_desc* D, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int L = D->L, d = D->d, M = D->M, S = D->S < 1 ? 1 : D->S;
  if (L < 1 || L > ST_MAX_LAYERS || d > kMaxD || padded(M) > MAX_MP || D->B < 1) return -2;
  const int MP = padded(M);
  const StepLayout y = step_layout(L, d, M, S, D->B);
  double* ws = D->workspace;
  const double kl_scale = (double)D->B / (double)D->num_data;
  unsigned int* clamp = reinterpret_cast<unsigned int*>(ws + y.clamp);

  int kinds[ST_MAX_LAYERS];
  const double* Zx[ST_MAX_LAYERS]; const double* zf[ST_MAX_LAYERS]; const double* theta[ST_MAX_LAYERS];
  const double* m[ST_MAX_LAYERS]; const double* Lq[ST_MAX_LAYERS];
  double* ops[ST_MAX_LAYERS]; const double* cops[ST_MAX_LAYERS]; const double* gops[ST_MAX_LAYERS];
  double* pre_work[ST_MAX_LAYERS]; double* dtheta_pre[ST_MAX_LAYERS]; double* dzf_pre[ST_MAX_LAYERS];
  double* dm_pre[ST_MAX_LAYERS]; double* dLq[ST_MAX_LAYERS];
  for (int l = 0; l < L; ++l) {
    const mobo_layer_desc& ld = D->layer[l];
    kinds[l] = l == 0 ? 0 : 1;
    Zx[l] = ld.Zx; zf[l] = l == 0 ? nullptr : D->layer[l - 1].m; theta[l] = ws + y.theta[l];
    m[l] = ld.m; Lq[l] = ld.Lq;
    ops[l] = ws + y.ops[l]; cops[l] = ops[l]; gops[l] = ws + y.gops[l];
    pre_work[l] = ws + y.pre_work[l]; dtheta_pre[l] = ws + y.dtheta_pre[l]; dzf_pre[l] = ws + y.dzf_pre[l];
    dm_pre[l] = ws + y.dm_pre[l];
    dLq[l] = (ld.g_Lq && !D->accumulate) ? ld.g_Lq : ws + y.dLq_tmp[l];
  }
  // 1. raw -> constrained
  {
    PrepArgs a;
    a.L = L; a.d = d; a.noise = ws + y.noise; a.clamp_count = clamp; a.dkl = kl_scale;
    for (int l = 0; l < ST_MAX_LAYERS; ++l) {
      const bool ok = l < L;
      for (int i = 0; i < ST_MAX_THETA; ++i) a.raw_theta[l][i] = ok ? D->layer[l].raw_theta[i] : nullptr;
      a.raw_noise[l] = ok ? D->layer[l].raw_noise : nullptr;
      a.noise_lo[l] = ok ? D->layer[l].noise_lower : 0.0; a.noise_hi[l] = ok ? D->layer[l].noise_upper : 0.0;
      a.theta[l] = ok ? ws + y.theta[l] : nullptr;
      a.gops_scal[l] = ok ? ws + y.gops[l] + ops_scal(MP) : nullptr;
      a.ops_scal[l] = ok ? ws + y.ops[l] + ops_scal(MP) : nullptr;
      a.oc_flags[l] = ok ? reinterpret_cast<int*>(ws + y.ops[l] + ops_flags(MP)) : nullptr;
    }
    MOBO_LAUNCH("step_prep_kernel", st, step_prep_kernel<<<L, 96, 0, st>>>(a));
  }
  // 2. operator chain of every layer
  MOBO_TRY(model_precompute(L, kinds, d, M, Zx, zf, theta, m, Lq, D->jitter, ops, stream, false));
  // 3. forward row passes, low -> high fidelity
  for (int l = 0; l < L; ++l) {
    const long long R = y.R[l];
    MOBO_TRY(mobo_layer_rows_fwd(kinds[l], d, M, Zx[l], zf[l], theta[l], ops[l], D->x, l == 0 ? 1 : S,
                                 l == 0 ? nullptr : ws + y.mu[l - 1], l == 0 ? nullptr : ws + y.var[l - 1],
                                 l == 0 ? 1 : (int)(R / y.R[l - 1]), D->layer[l].eps, R, nullptr, R, 1,
                                 ws + y.mu[l], ws + y.var[l], ws + y.craw[l], clamp + l, ws + y.Ts[l], ws + y.Us[l],
                                 stream));
  }
  // 4. ELBO terms and backward row passes, high -> low fidelity.  Per layer the main stream runs the SYRK statistics,
  //    the product and the covariance-gradient kernels; that layer's operator-chain backward (a dozen latency-bound
  //    M x M launches) runs on the side stream, hidden behind the row kernels.
  SideCtx* scp = D->ctx ? static_cast<SideCtx*>(D->ctx) : (g_side_stream_on ? default_side_ctx() : nullptr);
  const bool fork = g_side_stream_on && scp != nullptr;
  SideCtx dummy;
  SideCtx& sc = scp ? *scp : dummy;
  cudaStream_t ss = fork ? sc.s : st;
  for (int l = L - 1; l >= 0; --l) {
    const long long R = y.R[l];
    EllArgs e;
    e.layer = l; e.R = R; e.S = l == 0 ? 1 : S; e.y = D->y; e.fid = D->fid;
    e.mu = ws + y.mu[l]; e.var = ws + y.var[l]; e.noise = ws + y.noise;
    e.df_next = l + 1 < L ? ws + y.df[l + 1] : nullptr;
    e.eps_next = l + 1 < L ? D->layer[l + 1].eps : nullptr;
    e.prep_next = l + 1 < L ? (int)(y.R[l + 1] / R) : 1;
    e.dmu = ws + y.dmu[l]; e.dvar = ws + y.dvar[l]; e.part = ws + y.ell_part[l];
    MOBO_LAUNCH("ell_kernel", st, ell_kernel<<<y.ell_blocks[l], ELL_THREADS, 0, st>>>(e));
    // SYRK proper on the main stream (DMMA-bound like the row kernels: running them side by side only made both
    // slower); the fold of its partials and the latency-bound operator-chain backward of this layer go to the side
    // stream
    double* gl = ws + y.gops[l];
    MOBO_TRY(launch_syrk_both(ws + y.Ts[l], ws + y.dvar[l], ws + y.craw[l], MP, R, ws + y.stats0[l], ws + y.stats1[l],
                              clamp + l, ws + y.dmu[l], ws + y.stats_alpha[l], st));
    if (fork && (cudaEventRecord(sc.fork[l], st) != cudaSuccess || cudaStreamWaitEvent(ss, sc.fork[l], 0) != cudaSuccess))
      return -1;
    MOBO_TRY(launch_syrk_reduce(0, MP, R, ws + y.stats0[l], gl + ops_block(MP, OPS_W), clamp + l, ws + y.stats_alpha[l],
                                gl + ops_alpha(MP), nullptr, ss));
    MOBO_TRY(launch_syrk_reduce(1, MP, R, ws + y.stats1[l], gl + ops_block(MP, OPS_H), clamp + l, nullptr, nullptr,
                                gl + ops_scal(MP) + SC_CLAMP, ss));
    MOBO_TRY(mobo_model_precompute_bwd(1, kinds + l, d, M, Zx + l, zf + l, theta + l, m + l, Lq + l, cops + l, gops + l,
                                       pre_work + l, dtheta_pre + l, dzf_pre + l, dm_pre + l, dLq + l, (void*)ss));
    // d L_q of this layer is final (it is written straight into the caller's gradient): a multi-GPU caller may start
    // its all-reduce now, behind the lower layers' row kernels (mobo_step_ctx_wait_layer)
    if (scp && !D->accumulate && cudaEventRecord(sc.done[l], ss) != cudaSuccess) return -1;
    // main stream: dk = W^T dt and the covariance gradient
    RowArgs a;
    fill_row_args(a, kinds[l], d, M, Zx[l], zf[l], theta[l], ops[l], D->x, l == 0 ? 1 : S,
                  l == 0 ? nullptr : ws + y.mu[l - 1], l == 0 ? nullptr : ws + y.var[l - 1],
                  l == 0 ? 1 : (int)(R / y.R[l - 1]), D->layer[l].eps, R, nullptr, R, 1);
    fill_row_bwd_args(a, ws + y.dmu[l], ws + y.dvar[l], ws + y.craw[l], ws + y.Ts[l], ws + y.Us[l], 1, ws + y.df[l],
                      nullptr);
    a.clamp_count = clamp + l;
    a.sm_reserve = fork ? side_stream_sms() : 0;
    MOBO_TRY(rows_bwd_main(a, kinds[l], d, M, R, ws + y.rows_work, ws + y.kg_part[l], ws + y.dtheta_rows[l],
                           l == 0 ? nullptr : ws + y.dzf_rows[l], st, fork ? ss : nullptr, fork ? sc.kg[l] : nullptr));
  }
  // 5. join: the gradient assembly needs both streams' results
  if (fork && (cudaEventRecord(sc.join, ss) != cudaSuccess || cudaStreamWaitEvent(st, sc.join, 0) != cudaSuccess)) return -1;
  // 6. gradient assembly
  {
    FinishArgs a;
    a.L = L; a.d = d; a.M = M; a.MP = MP; a.kl_scale = kl_scale; a.out = D->out; a.accumulate = D->accumulate;
    for (int l = 0; l < ST_MAX_LAYERS; ++l) {
      const bool ok = l < L;
      for (int i = 0; i < ST_MAX_THETA; ++i) {
        a.raw_theta[l][i] = ok ? D->layer[l].raw_theta[i] : nullptr;
        a.g_raw_theta[l][i] = ok ? D->layer[l].g_raw_theta[i] : nullptr;
      }
      a.dtheta_rows[l] = ok ? ws + y.dtheta_rows[l] : nullptr;
      a.dtheta_pre[l] = ok ? ws + y.dtheta_pre[l] : nullptr;
      a.dzf_rows[l] = (ok && l > 0) ? ws + y.dzf_rows[l] : nullptr;
      a.dzf_pre[l] = (ok && l > 0) ? ws + y.dzf_pre[l] : nullptr;
      a.dm_pre[l] = ok ? ws + y.dm_pre[l] : nullptr;
      a.g_m[l] = ok ? D->layer[l].g_m : nullptr;
      a.raw_noise[l] = ok ? D->layer[l].raw_noise : nullptr;
      a.noise_lo[l] = ok ? D->layer[l].noise_lower : 0.0; a.noise_hi[l] = ok ? D->layer[l].noise_upper : 0.0;
      a.g_raw_noise[l] = ok ? D->layer[l].g_raw_noise : nullptr;
      a.ell_part[l] = ok ? ws + y.ell_part[l] : nullptr;
      a.ell_blocks[l] = ok ? y.ell_blocks[l] : 0;
      a.ops_scal[l] = ok ? ws + y.ops[l] + ops_scal(MP) : nullptr;
    }
    MOBO_LAUNCH("step_finish_kernel", st, step_finish_kernel<<<1, 256, 0, st>>>(a));
  }
  if (D->accumulate)
    for (int l = 0; l < L; ++l)
      if (D->layer[l].g_Lq) {
        const long long n = (long long)M * M;
        MOBO_LAUNCH("add_matrix_kernel", st,
                    add_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(D->layer[l].g_Lq, dLq[l], n));
      }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mobo_adam(int nt, const mobo_adam_tensor* tensors, double lr, double beta1, double beta2, double eps,
              long long step, long long* step_dev, const double* skip_flag, void* stream) {
  if (nt < 1 || nt > ADAM_MAX_TENSORS || (step < 1 && !step_dev)) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  AdamArgs a;
  a.nt = nt; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.step_dev = step_dev; a.skip = skip_flag;
  a.bc1 = step_dev ? 1.0 : 1.0 - pow(beta1, (double)step);
  a.bc2_sqrt = step_dev ? 1.0 : sqrt(1.0 - pow(beta2, (double)step));
  long long mx = 1;
  for (int i = 0; i < ADAM_MAX_TENSORS; ++i) {
    const bool ok = i < nt;
    a.p[i] = ok ? tensors[i].p : nullptr; a.g[i] = ok ? tensors[i].g : nullptr;
    a.m[i] = ok ? tensors[i].exp_avg : nullptr; a.v[i] = ok ? tensors[i].exp_avg_sq : nullptr;
    a.n[i] = ok ? tensors[i].n : 0;
    if (ok && tensors[i].n > mx) mx = tensors[i].n;
  }
  long long gx = (mx + 255) / 256;
  if (gx > 64) gx = 64;
  MOBO_LAUNCH("adam_kernel", st, adam_kernel<<<dim3((unsigned)gx, nt), 256, 0, st>>>(a));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mobo_adam_tick(long long* step_dev, const double* skip_flag, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MOBO_LAUNCH("adam_tick_kernel", st, adam_tick_kernel<<<1, 1, 0, st>>>(step_dev, skip_flag));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mobo_acq_moments(int fidelity, int d, int M, int S, long long n, const double* Zx, const double* const* zf,
                     const double* const* theta, const double* const* ops, const double* const* samples,
                     const double* raw_noise, double noise_lower, double noise_upper, const double* X,
                     double* out_mu, double* out_var, double* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (fidelity < 0 || fidelity >= ST_MAX_LAYERS || n < 1 || S < 1) return -2;
  const long long RS = n * (long long)S;
  double* mu[2] = {scratch, scratch + 2 * RS};
  double* var[2] = {scratch + RS, scratch + 3 * RS};
  long long Rprev = n;
  for (int l = 0; l <= fidelity; ++l) {
    const long long R = l == 0 ? n : RS;
    const int cur = l & 1, prv = cur ^ 1;
    MOBO_TRY(mobo_layer_rows_fwd(l == 0 ? 0 : 1, d, M, Zx, l == 0 ? nullptr : zf[l], theta[l], ops[l], X,
                                 l == 0 ? 1 : S, l == 0 ? nullptr : mu[prv], l == 0 ? nullptr : var[prv],
                                 l == 0 ? 1 : (int)(R / Rprev), l == 0 ? nullptr : samples[l], S, nullptr, R, 0,
                                 mu[cur], var[cur], nullptr, nullptr, nullptr, nullptr, stream));
    Rprev = R;
  }
  MomentArgs a;
  a.n = n; a.S = S; a.tiled = fidelity >= 1 ? 1 : 0;
  a.mu = mu[fidelity & 1]; a.var = var[fidelity & 1];
  a.noise_lo = noise_lower; a.noise_hi = noise_upper; a.raw_noise = raw_noise;
  a.out_mu = out_mu; a.out_var = out_var;
  MOBO_LAUNCH("moment_kernel", st, moment_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mobo_jes(const double* var_uncond, const double* var_cond, long long n, int accumulate, double* out,
             void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n < 1) return 0;
  MOBO_LAUNCH("jes_kernel", st,
              jes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(var_uncond, var_cond, out, n, accumulate));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mobo_rff_eval(int L, int d, int F, const double* const* params, const double* scales, const double* x, long long n,
                  double* f, double* grad, void* stream) {
  if (L < 1 || L > RFF_MAX_LAYERS) return -2;
  RffArgs a;
  a.L = L; a.d = d; a.F = F; a.want_grad = grad != nullptr;
  for (int l = 0; l < RFF_MAX_LAYERS; ++l) {
    a.params[l] = l < L ? params[l] : nullptr;
    for (int q = 0; q < 3; ++q) a.scale[l][q] = l < L ? scales[3 * l + q] : 0.0;
  }
  a.x = x; a.n = n; a.f = f; a.grad = grad;
  return launch_rff_eval(a, (cudaStream_t)stream);
}

int mobo_pareto_mask(const double* pts, long long n, int k, unsigned char* mask, void* stream) {
  return launch_pareto_mask(pts, n, k, mask, (cudaStream_t)stream);
}

}  // extern "C"
