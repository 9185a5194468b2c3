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
  return (size_t)6 * MP * MP + 2 * MP + (size_t)((M + KZB_WARPS - 1) / KZB_WARPS) * (5 + 2 * kMaxD) + 64;
}

int mobo_kzz(int kind, int d, int M, const double* Zx, const double* zf, const double* theta, double jitter,
             double* P, void* stream) {
  const int MP = padded(M);
  if (d > kMaxD || MP > MAX_MP) return -2;
  MOBO_LAUNCH("kzz_kernel", (cudaStream_t)stream, kzz_kernel<<<(MP * MP + 255) / 256, 256, 0, (cudaStream_t)stream>>>(kind, d, M, MP, Zx, zf, theta, jitter, P));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mobo_layer_precompute(int kind, int d, int M, const double* Zx, const double* zf, const double* theta,
                          const double* m, const double* Lq, double jitter, double* ops, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int MP = padded(M);
  if (d > kMaxD || MP > MAX_MP) return -2;
  double* L = ops + ops_block(MP, OPS_L);
  double* W = ops + ops_block(MP, OPS_W);
  double* WT = ops + ops_block(MP, OPS_WT);
  double* H = ops + ops_block(MP, OPS_H);
  double* HT = ops + ops_block(MP, OPS_HT);
  double* P = ops + ops_block(MP, OPS_P);
  double* LQ = ops + ops_block(MP, OPS_LQ);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_done = true;
  }
  MOBO_TRY(mobo_kzz(kind, d, M, Zx, zf, theta, jitter, P, stream));
  const size_t chol_smem = (size_t)(CH_NB * (CH_NB + 1) + 8 + (size_t)(MP - CH_NB) * CH_LDP) * sizeof(double);
  MOBO_LAUNCH("chol_kernel", st, chol_kernel<<<1, CH_THREADS, chol_smem, st>>>(P, L, MP, ops + ops_scal(MP)));
  const size_t tri_smem = (size_t)(MP / 32) * 32 * 36 * sizeof(double);
  MOBO_LAUNCH("trtri_kernel", st, trtri_kernel<<<1, TI_THREADS, tri_smem, st>>>(L, W, MP));
  MOBO_TRY(ew(EW_TRANSPOSE, M, MP, W, nullptr, WT, nullptr, 1.0, nullptr, nullptr, st));
  MOBO_TRY(ew(EW_PAD_TRIL, M, MP, Lq, nullptr, LQ, nullptr, 1.0, nullptr, nullptr, st));
  MOBO_TRY(gemm(MP, W, false, LQ, false, H, 1.0, 0.0, st));
  MOBO_TRY(ew(EW_TRANSPOSE, M, MP, H, nullptr, HT, nullptr, 1.0, nullptr, nullptr, st));
  MOBO_LAUNCH("finalize_kernel", st, finalize_kernel<<<1, 256, 0, st>>>(M, MP, m, ops));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int mobo_layer_precompute_bwd(int kind, int d, int M, const double* Zx, const double* zf, const double* theta,
                              const double* m, const double* Lq, const double* ops, const double* gops,
                              double* work, double* dtheta, double* dzf, double* dm, double* dLq, void* stream) {
  (void)Lq;
  cudaStream_t st = (cudaStream_t)stream;
  const int MP = padded(M);
  if (d > kMaxD || MP > MAX_MP) return -2;
  const size_t MP2 = (size_t)MP * MP;
  const double* L = ops + ops_block(MP, OPS_L);
  const double* W = ops + ops_block(MP, OPS_W);
  const double* WT = ops + ops_block(MP, OPS_WT);
  const double* H = ops + ops_block(MP, OPS_H);
  const double* HT = ops + ops_block(MP, OPS_HT);
  const double* LQ = ops + ops_block(MP, OPS_LQ);
  const double* beta = ops + ops_beta(MP);
  const double* A2 = gops + ops_block(MP, OPS_W);
  const double* Ac = gops + ops_block(MP, OPS_H);
  const double* dalpha = gops + ops_alpha(MP);
  const double* dkl = gops + ops_scal(MP) + SC_KL;
  double* T1 = work;
  double* T2 = work + MP2;
  double* T3 = work + 2 * MP2;
  double* T4 = work + 3 * MP2;   // dH
  double* T5 = work + 4 * MP2;   // dW
  double* T6 = work + 5 * MP2;
  double* dbeta = work + 6 * MP2;
  double* mpad = dbeta + MP;     // m zero-padded to MP
  double* part = mpad + MP;
  // A1 = A2 - Ac  (gradient of the clamped first quadratic form)
  MOBO_TRY(ew(EW_SUB, M, MP, A2, Ac, T1, nullptr, 1.0, nullptr, nullptr, st));
  // Y = H^T W ; dY = 2 Y A2
  MOBO_TRY(gemm(MP, HT, false, W, false, T2, 1.0, 0.0, st));
  MOBO_TRY(gemm(MP, T2, false, A2, false, T3, 2.0, 0.0, st));
  // dH = W dY^T + dkl H
  MOBO_TRY(ew(EW_SCALE, M, MP, H, nullptr, T4, dkl, 1.0, nullptr, nullptr, st));
  MOBO_TRY(gemm(MP, W, false, T3, true, T4, 1.0, 1.0, st));
  // dbeta = W dalpha + dkl beta ; dm = W^T dbeta
  MOBO_LAUNCH("dbeta_kernel", st, dbeta_kernel<<<1, 256, 0, st>>>(M, MP, ops, dalpha, dkl, dbeta, dm));
  cudaMemsetAsync(mpad, 0, sizeof(double) * MP, st);
  cudaMemcpyAsync(mpad, m, sizeof(double) * M, cudaMemcpyDeviceToDevice, st);
  // dW = -2 W A1 + H dY + beta dalpha^T + dbeta m^T + dH Lq^T   (lower triangle kept)
  MOBO_TRY(gemm(MP, W, false, T1, false, T5, -2.0, 0.0, st));
  MOBO_TRY(gemm(MP, H, false, T3, false, T5, 1.0, 1.0, st));
  MOBO_TRY(ew(EW_RANK1_ADD, M, MP, nullptr, nullptr, T5, nullptr, 1.0, beta, dalpha, st));
  MOBO_TRY(ew(EW_RANK1_ADD, M, MP, nullptr, nullptr, T5, nullptr, 1.0, dbeta, mpad, st));
  MOBO_TRY(gemm(MP, T4, false, LQ, true, T5, 1.0, 1.0, st));
  MOBO_TRY(ew(EW_TRIL_INPLACE, M, MP, T5, nullptr, T5, nullptr, 1.0, nullptr, nullptr, st));
  // dLq = tril(W^T dH) - dkl diag(1 / Lq_ii)
  MOBO_TRY(gemm(MP, WT, false, T4, false, T6, 1.0, 0.0, st));
  MOBO_LAUNCH("dlq_extract_kernel", st, dlq_extract_kernel<<<(M * M + 255) / 256, 256, 0, st>>>(M, MP, T6, LQ, dkl, dLq));
  // dL = -tril(W^T dW W^T) + dkl diag(1 / L_ii)
  MOBO_TRY(gemm(MP, WT, false, T5, false, T1, 1.0, 0.0, st));
  MOBO_TRY(gemm(MP, T1, false, WT, false, T2, 1.0, 0.0, st));
  MOBO_TRY(ew(EW_NEG_TRIL_DIAG, M, MP, T2, L, T3, dkl, 1.0, nullptr, nullptr, st));
  // dP = sym( W^T Phi(L^T dL) W )
  MOBO_TRY(gemm(MP, L, true, T3, false, T1, 1.0, 0.0, st));
  MOBO_TRY(ew(EW_PHI, M, MP, T1, nullptr, T2, nullptr, 1.0, nullptr, nullptr, st));
  MOBO_TRY(gemm(MP, WT, false, T2, false, T1, 1.0, 0.0, st));
  MOBO_TRY(gemm(MP, T1, false, W, false, T2, 1.0, 0.0, st));
  MOBO_TRY(ew(EW_SYM, M, MP, T2, nullptr, T3, nullptr, 1.0, nullptr, nullptr, st));
  // through K(Z, Z)
  const int nblk = (M + KZB_WARPS - 1) / KZB_WARPS;
  MOBO_LAUNCH("kzz_bwd_kernel", st, kzz_bwd_kernel<<<nblk, KZB_WARPS * 32, 0, st>>>(kind, d, M, MP, Zx, zf, theta, T3, part, dzf, 0));
  MOBO_TRY(launch_reduce_partials(part, nblk, theta_size(kind, d), 5 + 2 * kMaxD, dtheta, 0, st));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
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
  a.Ksave = nullptr; a.Tsave = nullptr; a.Usave = nullptr;
  a.dmu = nullptr; a.dvar = nullptr; a.df = nullptr; a.dxrow = nullptr; a.part_theta = nullptr; a.part_zf = nullptr;
  a.want_param_grads = 0; a.want_x_grads = 0;
}

int mobo_layer_rows_fwd(int kind, int d, int M, const double* Zx, const double* zf, const double* theta,
                        const double* ops, const double* x, int xrep, const double* mu_prev, const double* var_prev,
                        int prep, const double* eps, long long eps_mod, const double* f_direct, long long R,
                        int training, double* mu, double* var, double* craw, unsigned int* clamp_count,
                        double* Ksave, double* Tsave, double* Usave, void* stream) {
  RowArgs a;
  fill_row_args(a, kind, d, M, Zx, zf, theta, ops, x, xrep, mu_prev, var_prev, prep, eps, eps_mod, f_direct, R,
                training);
  a.mu = mu; a.var = var; a.craw = craw; a.clamp_count = clamp_count;
  a.Ksave = Ksave; a.Tsave = Tsave; a.Usave = Usave;
  return launch_row_fwd(a, (cudaStream_t)stream);
}

int mobo_layer_rows_bwd(int kind, int d, int M, const double* Zx, const double* zf, const double* theta,
                        const double* ops, const double* x, int xrep, const double* mu_prev, const double* var_prev,
                        int prep, const double* eps, long long eps_mod, const double* f_direct, long long R,
                        int training, const double* dmu, const double* dvar, const double* craw,
                        const unsigned int* clamp_count, const double* Ksave, const double* Tsave,
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
    MOBO_TRY(launch_syrk(Ksave, dvar, craw, 0, MP, R, part_syrk, gops + ops_block(MP, OPS_W), clamp_count, dmu,
                         part_alpha, gops + ops_alpha(MP), st));
    MOBO_TRY(launch_syrk(Ksave, dvar, craw, 1, MP, R, part_syrk, gops + ops_block(MP, OPS_H),
                         training ? clamp_count : nullptr, nullptr, nullptr, nullptr, st));
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
