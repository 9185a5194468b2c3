// extern "C" entry points of libmobocmf_b200.so (declared in include/mobocmf_b200.h) and the host-side kernel
// sequences behind them.
#include <string.h>
#include "../../include/mobocmf_b200.h"
#include "matrix_ops.cu"
#include "row_pass.cu"

using namespace mobo;

static inline int padded(int M) { return ((M + 31) / 32) * 32; }
#define MOBO_TRY(x) do { int _e = (x); if (_e) return _e; } while (0)

extern "C" {

int mobo_abi_version(void) { return 100; }

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
  return (size_t)148 * 2 * (MAX_THETA + MP) + syrk_part_doubles(MP, R) + (size_t)syrk_nchunk(MP, R) * MP + 64;
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
  MOBO_LAUNCH("kzz_kernel", st, kzz_kernel<<<dim3((MP * MP + 255) / 256, 1), 256, 0, st>>>(b, jitter, OPS_P));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mobo_model_precompute(int nl, const int* kinds, int d, int M, const double* const* Zx, const double* const* zf,
                          const double* const* theta, const double* const* m, const double* const* Lq,
                          double jitter, double* const* ops, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LayerBatch b;
  MOBO_TRY(fill_batch(b, nl, kinds, d, M, Zx, zf, theta, m, Lq, ops));
  const int MP = b.MP;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_done = true;
  }
  const dim3 ew_grid((MP * MP + 255) / 256, nl);
  MOBO_LAUNCH("kzz_kernel", st, kzz_kernel<<<ew_grid, 256, 0, st>>>(b, jitter, OPS_P));
  const size_t chol_smem = (size_t)(64 * CH_LD + (size_t)MP * CH_LD) * sizeof(double);
  MOBO_LAUNCH("chol_inv_kernel", st, chol_inv_kernel<<<nl, CH_THREADS, chol_smem, st>>>(b));
  MOBO_LAUNCH("padtril_kernel", st, padtril_kernel<<<ew_grid, 256, 0, st>>>(b));
  GemmOperand A, B;
  double* C[MAX_BATCH]; double* Ct[MAX_BATCH];
  FinBatch f;
  for (int i = 0; i < MAX_BATCH; ++i) {
    const bool ok = i < nl;
    A.p[i] = ok ? ops[i] + ops_block(MP, OPS_W) : nullptr;
    B.p[i] = ok ? ops[i] + ops_block(MP, OPS_LQ) : nullptr;
    C[i] = ok ? ops[i] + ops_block(MP, OPS_H) : nullptr;
    Ct[i] = ok ? ops[i] + ops_block(MP, OPS_HT) : nullptr;
    f.rowstat[i] = ok ? ops[i] + ops_rowstat(MP) : nullptr;
    f.counter[i] = ok ? reinterpret_cast<unsigned int*>(ops[i] + ops_scal(MP) + SC_COUNTER) : nullptr;
  }
  A.trans = false; A.tri = 1; B.trans = false; B.tri = 1;
  MOBO_TRY(gemm_batched(nl, MP, A, B, C, Ct, 1.0, 0.0, nullptr, st));
  MOBO_LAUNCH("finalize_kernel", st,
              finalize_kernel<<<dim3((MP + FIN_WARPS - 1) / FIN_WARPS, nl), FIN_WARPS * 32, 0, st>>>(b, f));
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
  a.dmu = nullptr; a.dvar = nullptr; a.df = nullptr; a.dxrow = nullptr; a.part_theta = nullptr; a.part_zf = nullptr;
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
  const int MP = a.MP;
  a.dmu = dmu; a.dvar = dvar; a.craw = const_cast<double*>(craw);
  a.Tsave = const_cast<double*>(Tsave); a.Usave = const_cast<double*>(Usave);
  a.df = df; a.dxrow = dxrow;
  a.want_param_grads = want_param_grads; a.want_x_grads = dxrow != nullptr;
  if (!a.want_param_grads && !a.want_x_grads) a.want_param_grads = 1;
  const int grid = row_grid(R);
  double* part_theta = work;
  double* part_zf = part_theta + (size_t)grid * MAX_THETA;
  double* part_syrk = part_zf + (size_t)grid * MP;
  double* part_alpha = part_syrk + syrk_part_doubles(MP, R);
  a.part_theta = part_theta; a.part_zf = part_zf;
  MOBO_TRY(launch_row_bwd(a, grid, st));
  if (a.want_param_grads && dtheta) {
    MOBO_TRY(launch_reduce_partials(part_theta, grid, theta_size(kind, d), MAX_THETA, dtheta, 0, st));
    if (kind == 1 && dzf) MOBO_TRY(launch_reduce_partials(part_zf, grid, M, MP, dzf, 0, st));
  }
  if (gops) {
    MOBO_TRY(launch_syrk(Tsave, dvar, craw, 0, MP, R, part_syrk, gops + ops_block(MP, OPS_W), clamp_count, dmu,
                         part_alpha, gops + ops_alpha(MP), nullptr, st));
    MOBO_TRY(launch_syrk(Tsave, dvar, craw, 1, MP, R, part_syrk, gops + ops_block(MP, OPS_H),
                         training ? clamp_count : nullptr, nullptr, nullptr, nullptr,
                         gops + ops_scal(MP) + SC_CLAMP, st));
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
