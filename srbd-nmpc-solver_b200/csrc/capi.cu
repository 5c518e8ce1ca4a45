// capi.cu — the C-ABI of include/srbd_b200.h over the sm_100a kernels.  No CPU fallback: without a
// usable CUDA device every compute entry point fails with SRBD_ERR_CUDA.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "ipm_solve.cuh"
#include "ipm_srbd.cuh"
#include "srbd_model.cuh"

using namespace srbd;

struct srbd_ctx {
  int device = 0, B = 0;
  srbd_qp_dims dims{};
  QpLayout L{};
  bool is_srbd = false;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  srbd_model_params model{};
  srbd_ipm_args args{};
  ModelDev* d_model = nullptr;
  // NMPC level
  double *d_x = nullptr, *d_u = nullptr, *d_xref = nullptr, *d_x0abs = nullptr, *d_defect = nullptr;
  double* d_srec = nullptr;  // [B][N+1][kSrec] compact stage records (K2 -> SRBD K3 variant)
  double* d_gdyn = nullptr;  // [B][N][kBabtDyn] stage-dependent chunks of the BAbt records (K1 -> SRBD K3 variant, layout.cuh)
  bool gdyn_valid = false;   // K1 wrote d_babt and d_gdyn last (not an upload / pack)
  double* d_gconst = nullptr;   // [336] dense BAbt record of (QP 0, stage 1): the model constants of K3's compact BAbt streaming
  bool babt_dense_valid = false;   // the dense BAbt records of the current linearization exist (else: dyn records only)
  long long lin_traj_version = -1; // trajectory version K1 last ran on
  uint8_t* d_contact = nullptr;
  bool have_contact = false;
  double* d_alpha = nullptr; int* d_conv = nullptr; double* d_merit = nullptr;
  // packed QP
  double *d_babt = nullptr, *d_rsq = nullptr, *d_dct = nullptr, *d_d = nullptr, *d_dmask = nullptr, *d_raw0 = nullptr;
  double* d_x0 = nullptr;  // [B][nx] QP-level initial state (delta form on the NMPC path)
  double* d_r0raw = nullptr;  // [B][nu] un-embedded r0 of uploaded QPs (closed-loop driver)
  // closed-loop MPC driver (lazy)
  double *d_mpc_x = nullptr, *d_mpc_u = nullptr, *d_mpc_xcur = nullptr, *d_plantA = nullptr, *d_plantB = nullptr, *d_plantb = nullptr;
  int *d_mpc_iter = nullptr, *d_mpc_status = nullptr;
  int mpc_steps_alloc = 0, plant_alloc = 0;
  bool warm_from_solution = false;  // the next solve takes the previous solution (on the device) as the primal warm start
  // outputs live in ONE device arena [x | u | pi | res_max | iter | status | lam | t] (srbd_out_layout): a caller with a
  // host arena of the same layout fetches them with a single D2H copy (srbd_download_packed)
  double* d_out = nullptr;
  size_t out_off[8] = {0, 0, 0, 0, 0, 0, 0, 0}, out_total = 0;
  // QP-level uploads whose host fields lie in one contiguous (pinned) range take ONE H2D copy into this arena
  char* d_in = nullptr;
  size_t in_bytes = 0;
  // uploaded QPs with the structure K2 guarantees (S = 0, C = 0, constant diagonal Q, one constant block D, upper side
  // masked) may take the tensor-core variant: detected on the device at upload
  ModelDev* d_model_qp = nullptr;
  int* d_flag = nullptr;
  int* h_flag = nullptr;     // pinned
  bool upload_variant_ok = false;
  // device-side SQP loop (srbd_sqp_solve): per-iteration counters of the QPs still iterating (gate of the next iteration's
  // kernels), per-QP iteration counts; sqp_loop: the launches below run gated / frozen
  int* d_active = nullptr;
  int* d_sqp_iter = nullptr;
  int active_alloc = 0;
  bool sqp_loop = false;
  const int* cur_gate = nullptr;
  // low-latency host -> host path (srbd_solve_host_graph): the whole pipeline as ONE CUDA graph launch over pinned staging
  cudaGraphExec_t graph_exec = nullptr;
  int graph_mode = -1, graph_contact = -1;
  long long graph_launches = 0;
  char* h_stage = nullptr;   // pinned: [x | u | xref | x0 | sol_x | sol_u | iter | status | contact]
  double *d_xinit = nullptr, *d_uinit = nullptr;
  bool have_init = false, packed = false, solved = false;
  // raw QP-level staging (lazy)
  std::vector<void*> raw_dev;
  srbd_qp_host d_qp{};
  bool raw_alloc = false;
  // outputs
  double *d_sol_x = nullptr, *d_sol_u = nullptr, *d_sol_pi = nullptr, *d_sol_lam = nullptr, *d_sol_t = nullptr;
  double *d_P = nullptr, *d_p = nullptr, *d_K = nullptr, *d_k = nullptr, *d_stat = nullptr;
  double* d_Lr0 = nullptr;   // [B][nu*nu] Cholesky factor of stage 0 (with the Riccati exports; srbd_download_ric_lr0)
  int *d_iter = nullptr, *d_status = nullptr, *d_counter = nullptr;
  double* d_resmax = nullptr;
  srbd_batch_stats* d_bstats = nullptr;
  double* d_ws = nullptr;
  double* d_ws2 = nullptr;   // workspace of the SRBD throughput variant of K3 (ipm_srbd.cuh)
  int* d_retry = nullptr;    // [B + 1]: rescue list of the variant (QPs that did not converge) and, at [B], its length
  int* d_retry2 = nullptr;   // [B + 1]: what the first rescue stage (the variant with the other pivot rounding) leaves over
  int grid2 = 0;
  int assembled_mode = -1;   // >= 0: the packed QP came from srbd_assemble (K2) in that mode
  int grid = 0, stat_rows = 0;
  bool spread = false;   // sparse batch: K3's kSpread instantiations on the full grid (solve_srbd_variant)
  bool upload_d_shared = false;   // srbd_qp_upload_layout: qp->D is ONE ng x nu matrix shared by every stage of every QP
  bool export_ric = false, export_stat = false;
  bool ric_valid = false, stat_valid = false;
  // K1 / K2 write the dense records (RSQrq, DCt, d, dmask, raw stage-0 blocks) only when something will read them: the
  // generic kernel, the getters, the rescue pass.  The throughput path (SRBD variant of K3) reads BAbt and the compact
  // stage records only; the dense ones are then materialised lazily from the unchanged trajectory (ensure_dense).
  bool dense_valid = false, raw0_valid = false;
  long long traj_version = 0, asm_traj_version = -1;  // the LAST solve wrote the Riccati exports (+ pi[0]) / the statistics table
  long long launches = 0;
  std::string err;
  int sm_count = 0;
  int smem_optin = 0;   // cudaDevAttrMaxSharedMemoryPerBlockOptin
};

namespace {

int fail(srbd_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, SRBD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)

template <class T>
cudaError_t dalloc(T** p, size_t n) {
  return cudaMalloc(reinterpret_cast<void**>(p), (n ? n : 1) * sizeof(T));
}

void fill_Ac(const srbd_model_params& m, double* Ac) {  // SRBD_model.cpp:244-255, row-major [g][j]
  std::memset(Ac, 0, sizeof(double) * 288);
  for (int leg = 0; leg < 2; ++leg) {
    const double* Rf = m.foot_rot + 9 * leg;
    const double *cx = Rf, *cy = Rf + 3, *cz = Rf + 6;
    auto A = [&](int i, int j) -> double& { return Ac[(12 * leg + i) * 12 + 6 * leg + j]; };
    A(0, 0) = -1; A(0, 2) = m.mu;
    A(1, 1) = -1; A(1, 2) = m.mu;
    A(2, 0) = 1;  A(2, 2) = m.mu;
    A(3, 1) = 1;  A(3, 2) = m.mu;
    A(4, 2) = -1;
    A(5, 2) = 1;
    for (int j = 0; j < 3; ++j) {
      A(6, j) = m.Lfx * cz[j];  A(6, 3 + j) = -cy[j];
      A(7, j) = m.Lfx * cz[j];  A(7, 3 + j) = cy[j];
      A(8, j) = m.Lfz * cz[j];  A(8, 3 + j) = -cz[j];
      A(9, j) = m.Lfz * cz[j];  A(9, 3 + j) = cz[j];
      A(10, 3 + j) = -cx[j];
      A(11, 3 + j) = cx[j];
    }
  }
}

typedef void (*ipm_kernel_t)(const IpmParams);
struct KernelChoice {
  ipm_kernel_t fn;
  const char* name;
};
using SrbdDims = SDims<12, 12, 0, 0, 24, 0>;
using QuadDims = SDims<12, 4, 3, 4, 0, 0>;
using Rnd0Dims = SDims<5, 3, 0, 0, 0, 0>;
using Rnd1Dims = SDims<5, 3, 2, 3, 2, 2>;

KernelChoice pick_kernel(const QpLayout& L) {
  if (SrbdDims::matches(L)) return {ipm_solve_kernel<SrbdDims>, "srbd"};
  if (QuadDims::matches(L)) return {ipm_solve_kernel<QuadDims>, "quadcopter"};
  if (Rnd0Dims::matches(L)) return {ipm_solve_kernel<Rnd0Dims>, "rnd0"};
  if (Rnd1Dims::matches(L)) return {ipm_solve_kernel<Rnd1Dims>, "rnd1"};
  return {ipm_solve_kernel<DDims>, "dynamic"};
}

int raw0_stride(const QpLayout& L) { return 2 * L.nx * L.nx + 2 * L.nx * L.nu + 2 * L.nx; }

// Do the dimensions and settings allow the SRBD tensor-core variant of K3 (cold start, classical Riccati, no Riccati /
// statistics exports)?
// exports_ok: the caller can hand the variant the raw stage-0 blocks its Riccati export needs (QP-level uploads: pack_kernel
// always writes them; K2 writes S0, Q0, q0 only with the dense records)
bool settings_allow_variant(const srbd_ctx* ctx, bool exports_ok = false) {
  const char* force = std::getenv("SRBD_K3_GENERIC");
  if (force && force[0] == '1') return false;
  if (!exports_ok && (ctx->export_ric || ctx->export_stat)) return false;
  return ctx->is_srbd && !ctx->args.warm_start && ctx->args.ric_alg == 0 &&
         ctx->args.itref_pred_max == 0 && ctx->args.itref_corr_max == 0;   // (the variant has no iterative refinement)
}
// Will srbd_qp_solve take the variant for a QP assembled by K2 under the current settings?
bool variant_eligible(const srbd_ctx* ctx) {
  if (!settings_allow_variant(ctx)) return false;
  // the variant relies on Ac being two 12x6 blocks (SRBD_model.cpp:244: Ac.block<12,6>(12*leg, 6*leg))
  double Ac[288];
  fill_Ac(ctx->model, Ac);
  for (int g = 0; g < 24; ++g)
    for (int j = 0; j < 12; ++j)
      if (Ac[g * 12 + j] != 0.0 && (j / 6) != (g / 12)) return false;
  return true;
}
bool want_dense(const srbd_ctx* ctx) {
  const char* d = std::getenv("SRBD_K2_DENSE");   // diagnosis: always write the dense records
  return (d && d[0] == '1') || !variant_eligible(ctx);
}

}  // namespace

extern "C" {

void srbd_model_params_default(srbd_model_params* p, int horizon) {
  std::memset(p, 0, sizeof(*p));
  p->mass = 15.0;                                           // NMPC_solver.cpp:334
  p->dt = 0.015;                                            // config/mpc_option.yaml:5
  const double Ib[3] = {0.541667, 0.516667, 1.0416667};     // mpc_option.yaml:10 (stored inverted, SRBD_model.cpp:46-49)
  for (int i = 0; i < 3; ++i) p->inertia_inv[i + 3 * i] = 1.0 / Ib[i];
  p->foot_pos[1] = -0.1;                                    // NMPC_solver.cpp:337
  p->foot_pos[4] = 0.1;
  for (int leg = 0; leg < 2; ++leg)
    for (int i = 0; i < 3; ++i) p->foot_rot[9 * leg + i + 3 * i] = 1.0;
  p->mu = 0.5; p->Lfx = 0.05; p->Lfz = 0.05; p->fmax = 1000.0; p->fmin = 0.0;  // SRBD_model.cpp:13-17
  p->gravity[2] = -9.8;                                     // SRBD_model.cpp:98
  p->Q[11] = 10.0;                                          // mpc_option.yaml:2
  const double Qf[12] = {0.5, 0.5, 0.5, 0.01, 0.01, 0.01, 100, 100, 100, 0.0, 0.0, 100};
  for (int i = 0; i < 12; ++i) p->Qf[i] = (double)horizon * Qf[i];  // NMPC_solver.cpp:58
  p->R = 1e-4;
  p->mu_b = 0.1; p->theta_b = 5.0;
  p->swing_fmax = 1.0;
}

void srbd_ipm_args_default(srbd_ipm_args* a) {
  std::memset(a, 0, sizeof(*a));
  a->iter_max = 15; a->alpha_min = 1e-8; a->mu0 = 1e2;     // ocp_qp_ipm_solver_settings.hpp:26-86
  a->tol_stat = a->tol_eq = a->tol_ineq = a->tol_comp = 1e-8;
  a->reg_prim = 1e-12; a->warm_start = 0; a->pred_corr = 1; a->ric_alg = 1; a->split_step = 0;
  a->cond_pred_corr = 1; a->cond_factor = 2.0; a->thr0 = 0.1;  // HPIPM SPEED-mode hidden defaults
  a->lam_min = a->t_min = a->tau_min = 1e-16; a->t_lam_min = 2; a->alpha_shorten = 1;
  a->itref_pred_max = 0; a->itref_corr_max = 0; a->itref_abs = 1.0; a->itref_rel = 1e-3;
}

int srbd_ipm_args_set_mode(srbd_ipm_args* a, int mode) {
  if (!a || mode < 0 || mode > 3) return SRBD_ERR_ARG;
  a->cond_pred_corr = mode == 0 ? 0 : 1;
  a->itref_pred_max = 0;
  a->itref_corr_max = mode == 2 ? 2 : (mode == 3 ? 4 : 0);
  return SRBD_OK;
}

size_t srbd_qp_nct(const srbd_qp_dims* d) {
  size_t n = 0;
  for (int k = 0; k <= d->N; ++k) {
    const int nb = (k < d->N ? d->nbu : 0) + (k > 0 ? d->nbx : 0);
    const int ng = k < d->N ? d->ng : d->ngN;
    n += 2 * (size_t)(nb + ng);
  }
  return n;
}

const char* srbd_last_error(const srbd_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int srbd_ctx_create(int device, int batch, const srbd_qp_dims* dims, void* stream, srbd_ctx** out) {
  if (!out || !dims || batch < 1) return SRBD_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) return SRBD_ERR_CUDA;
  srbd_ctx* ctx = new srbd_ctx();
  ctx->device = device;
  ctx->B = batch;
  ctx->dims = *dims;
  if (make_layout(*dims, nullptr, nullptr, &ctx->L) != 0) { delete ctx; return SRBD_ERR_ARG; }
  ctx->is_srbd = dims->nx == 12 && dims->nu == 12 && dims->ng == 24 && dims->nbx == 0 && dims->nbu == 0 && dims->ngN == 0;
  auto bail = [&](int code) { srbd_ctx_destroy(ctx); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return bail(SRBD_ERR_CUDA);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(SRBD_ERR_CUDA);
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (stream) ctx->stream = (cudaStream_t)stream;
  else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(SRBD_ERR_CUDA);
    ctx->own_stream = true;
  }
  const QpLayout& L = ctx->L;
  const size_t B = batch, S = L.N + 1, N = L.N;
  bool ok = true;
  auto A = [&](cudaError_t e) { ok = ok && (e == cudaSuccess); };
  A(dalloc(&ctx->d_model, 1));
  if (ctx->is_srbd) {
    A(dalloc(&ctx->d_x, B * S * 12)); A(dalloc(&ctx->d_u, B * N * 12)); A(dalloc(&ctx->d_xref, B * S * 12));
    A(dalloc(&ctx->d_x0abs, B * 12)); A(dalloc(&ctx->d_defect, B * N * 12)); A(dalloc(&ctx->d_contact, B * N * 2));
    A(dalloc(&ctx->d_alpha, B)); A(dalloc(&ctx->d_conv, B)); A(dalloc(&ctx->d_merit, B * 3));
    A(dalloc(&ctx->d_srec, B * S * kSrec));
    A(dalloc(&ctx->d_gdyn, B * N * kBabtDyn));
    A(dalloc(&ctx->d_gconst, 336));
  }
  A(dalloc(&ctx->d_babt, B * N * L.babt_stride)); A(dalloc(&ctx->d_rsq, B * S * L.rsq_stride));
  A(dalloc(&ctx->d_dct, B * S * L.dct_stride)); A(dalloc(&ctx->d_d, B * S * L.d_stride));
  A(dalloc(&ctx->d_dmask, B * S * L.d_stride)); A(dalloc(&ctx->d_raw0, B * raw0_stride(L)));
  A(dalloc(&ctx->d_x0, B * L.nx)); A(dalloc(&ctx->d_xinit, B * S * L.nx)); A(dalloc(&ctx->d_uinit, B * N * L.nu));
  A(dalloc(&ctx->d_r0raw, B * L.nu));
  {
    // output arena (offsets in doubles, every block on a 256-byte boundary)
    const size_t sizes[8] = {B * S * L.nx, B * N * L.nu, B * S * L.nx, B * 4, (B + 1) / 2, (B + 1) / 2,
                             B * (size_t)L.nct, B * (size_t)L.nct};
    size_t o = 0;
    for (int i = 0; i < 8; ++i) { ctx->out_off[i] = o; o += (sizes[i] + 31) & ~(size_t)31; }
    ctx->out_total = o;
    A(dalloc(&ctx->d_out, o));
    if (ok) {
      ctx->d_sol_x = ctx->d_out + ctx->out_off[0]; ctx->d_sol_u = ctx->d_out + ctx->out_off[1];
      ctx->d_sol_pi = ctx->d_out + ctx->out_off[2]; ctx->d_resmax = ctx->d_out + ctx->out_off[3];
      ctx->d_iter = reinterpret_cast<int*>(ctx->d_out + ctx->out_off[4]);
      ctx->d_status = reinterpret_cast<int*>(ctx->d_out + ctx->out_off[5]);
      ctx->d_sol_lam = ctx->d_out + ctx->out_off[6]; ctx->d_sol_t = ctx->d_out + ctx->out_off[7];
    }
  }
  A(dalloc(&ctx->d_counter, 1)); A(dalloc(&ctx->d_bstats, 1));
  A(dalloc(&ctx->d_model_qp, 1)); A(dalloc(&ctx->d_flag, 1));
  ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_flag), sizeof(int), cudaHostAllocDefault) == cudaSuccess;
  if (!ok) return bail(SRBD_ERR_CUDA);
  // persistent grid: every resident warp gets a private workspace
  KernelChoice kc = pick_kernel(L);
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kc.fn, 32, 0) != cudaSuccess || occ < 1) return bail(SRBD_ERR_CUDA);
  long long g = (long long)occ * ctx->sm_count;
  if (g > batch) g = batch;
  ctx->grid = (int)g;
  if (dalloc(&ctx->d_ws, (size_t)ctx->grid * L.ws_size) != cudaSuccess) return bail(SRBD_ERR_CUDA);
  if (cudaMemsetAsync(ctx->d_sol_pi, 0, B * S * L.nx * sizeof(double), ctx->stream) != cudaSuccess) return bail(SRBD_ERR_CUDA);
  srbd_model_params mp;
  srbd_model_params_default(&mp, L.N);
  srbd_ipm_args ia;
  srbd_ipm_args_default(&ia);
  ctx->args = ia;
  ctx->stat_rows = ia.iter_max + 2;
  if (srbd_set_model(ctx, &mp) != 0) return bail(SRBD_ERR_CUDA);
  if (ctx->is_srbd) {
    std::vector<double> ones(B, 1.0);
    if (cudaMemcpyAsync(ctx->d_alpha, ones.data(), B * sizeof(double), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        cudaMemsetAsync(ctx->d_conv, 0, B * sizeof(int), ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      return bail(SRBD_ERR_CUDA);
  }
  *out = ctx;
  return SRBD_OK;
}

int srbd_ctx_destroy(srbd_ctx* ctx) {
  if (!ctx) return SRBD_OK;
  cudaSetDevice(ctx->device);
  void* ptrs[] = {ctx->d_model, ctx->d_x, ctx->d_u, ctx->d_xref, ctx->d_x0abs, ctx->d_defect, ctx->d_contact,
                  ctx->d_alpha, ctx->d_conv, ctx->d_merit, ctx->d_babt, ctx->d_rsq, ctx->d_dct, ctx->d_d,
                  ctx->d_dmask, ctx->d_raw0, ctx->d_x0, ctx->d_xinit, ctx->d_uinit, ctx->d_out, ctx->d_in, ctx->d_model_qp,
                  ctx->d_flag, ctx->d_P, ctx->d_p, ctx->d_K, ctx->d_k, ctx->d_Lr0,
                  ctx->d_stat, ctx->d_counter, ctx->d_bstats, ctx->d_ws,
                  ctx->d_ws2, ctx->d_srec, ctx->d_gdyn, ctx->d_gconst, ctx->d_retry, ctx->d_retry2, ctx->d_active, ctx->d_sqp_iter, ctx->d_r0raw, ctx->d_mpc_x, ctx->d_mpc_u, ctx->d_mpc_xcur,
                  ctx->d_plantA, ctx->d_plantB, ctx->d_plantb, ctx->d_mpc_iter, ctx->d_mpc_status};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (void* p : ctx->raw_dev)
    if (p) cudaFree(p);
  if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
  if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SRBD_OK;
}

int srbd_set_model(srbd_ctx* ctx, const srbd_model_params* p) {
  if (!ctx || !p) return SRBD_ERR_ARG;
  ctx->model = *p;
  ModelDev md;
  md.m = *p;
  fill_Ac(*p, md.Ac);
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(ctx->d_model, &md, sizeof(md), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

int srbd_set_ipm_args(srbd_ctx* ctx, const srbd_ipm_args* a) {
  if (!ctx || !a) return SRBD_ERR_ARG;
  if (a->iter_max < 0 || a->iter_max > 1000) return fail(ctx, SRBD_ERR_ARG, "iter_max out of range");
  if (a->ric_alg != 0 && a->ric_alg != 1) return fail(ctx, SRBD_ERR_ARG, "ric_alg must be 0 (classical) or 1 (square root)");
  if (a->itref_pred_max < 0 || a->itref_corr_max < 0 || a->itref_pred_max > 16 || a->itref_corr_max > 16)
    return fail(ctx, SRBD_ERR_ARG, "itref_pred_max / itref_corr_max out of range");
  ctx->args = *a;
  const int rows = a->iter_max + 2;
  if (rows != ctx->stat_rows) {
    ctx->stat_rows = rows;
    if (ctx->d_stat) { cudaFree(ctx->d_stat); ctx->d_stat = nullptr; }
  }
  return SRBD_OK;
}

int srbd_set_outputs(srbd_ctx* ctx, int export_ric, int export_stat) {
  if (!ctx) return SRBD_ERR_ARG;
  ctx->export_ric = export_ric != 0;
  ctx->export_stat = export_stat != 0;
  return SRBD_OK;
}

int srbd_ctx_stat_rows(const srbd_ctx* ctx) { return ctx ? ctx->stat_rows : 0; }
void* srbd_ctx_stream(const srbd_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
long long srbd_ctx_launch_count(const srbd_ctx* ctx) { return ctx ? ctx->launches : 0; }

int srbd_ctx_sync(srbd_ctx* ctx) {
  if (!ctx) return SRBD_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

static int ensure_dense(srbd_ctx* ctx);
static int ensure_babt(srbd_ctx* ctx);

int srbd_ctx_device_ptr(srbd_ctx* ctx, int buf, void** ptr, size_t* bytes) {
  if (!ctx || !ptr || !bytes) return SRBD_ERR_ARG;
  const QpLayout& L = ctx->L;
  const size_t B = ctx->B, S = L.N + 1, N = L.N, D = sizeof(double);
  void* p = nullptr;
  size_t n = 0;
  switch (buf) {
    case SRBD_BUF_TRAJ_X: p = ctx->d_x; n = B * S * 12 * D; break;
    case SRBD_BUF_TRAJ_U: p = ctx->d_u; n = B * N * 12 * D; break;
    case SRBD_BUF_TRAJ_XREF: p = ctx->d_xref; n = B * S * 12 * D; break;
    case SRBD_BUF_X0: p = ctx->d_x0abs; n = B * 12 * D; break;
    case SRBD_BUF_CONTACT: p = ctx->d_contact; n = B * N * 2; break;
    case SRBD_BUF_SOL_X: p = ctx->d_sol_x; n = B * S * L.nx * D; break;
    case SRBD_BUF_SOL_U: p = ctx->d_sol_u; n = B * N * L.nu * D; break;
    case SRBD_BUF_SOL_PI: p = ctx->d_sol_pi; n = B * S * L.nx * D; break;
    case SRBD_BUF_SOL_LAM: p = ctx->d_sol_lam; n = B * (size_t)L.nct * D; break;
    case SRBD_BUF_SOL_T: p = ctx->d_sol_t; n = B * (size_t)L.nct * D; break;
    case SRBD_BUF_ITER: p = ctx->d_iter; n = B * sizeof(int); break;
    case SRBD_BUF_STATUS: p = ctx->d_status; n = B * sizeof(int); break;
    case SRBD_BUF_RESMAX: p = ctx->d_resmax; n = B * 4 * D; break;
    case SRBD_BUF_BABT: p = ctx->d_babt; n = B * N * L.babt_stride * D; break;
    case SRBD_BUF_RSQRQ: p = ctx->d_rsq; n = B * S * L.rsq_stride * D; break;
    case SRBD_BUF_DCT: p = ctx->d_dct; n = B * S * L.dct_stride * D; break;
    case SRBD_BUF_D: p = ctx->d_d; n = B * S * L.d_stride * D; break;
    case SRBD_BUF_DMASK: p = ctx->d_dmask; n = B * S * L.d_stride * D; break;
    case SRBD_BUF_DEFECT: p = ctx->d_defect; n = B * N * 12 * D; break;
    case SRBD_BUF_STAGE_REC: p = ctx->d_srec; n = ctx->d_srec ? B * S * kSrec * D : 0; break;
    case SRBD_BUF_BABT_DYN: p = ctx->gdyn_valid ? ctx->d_gdyn : nullptr; n = p ? B * N * kBabtDyn * D : 0; break;
    default: return fail(ctx, SRBD_ERR_ARG, "unknown buffer id");
  }
  if (buf == SRBD_BUF_BABT) {
    const long long l0 = ctx->launches;
    if (int rc = ensure_babt(ctx)) return rc;
    if (ctx->launches != l0) CU(cudaStreamSynchronize(ctx->stream));
  }
  if (buf == SRBD_BUF_RSQRQ || buf == SRBD_BUF_DCT || buf == SRBD_BUF_D || buf == SRBD_BUF_DMASK)
    if (ctx->packed) {
      const long long l0 = ctx->launches;
      if (int rc = ensure_dense(ctx)) return rc;
      // the caller reads the buffer on ITS stream: what was just launched on the context's stream must be done
      if (ctx->launches != l0) CU(cudaStreamSynchronize(ctx->stream));
    }
  if (!p) return fail(ctx, SRBD_ERR_STATE, "buffer not allocated for these dimensions");
  *ptr = p;
  *bytes = n;
  return SRBD_OK;
}

// ---- NMPC level ----------------------------------------------------------------------------------
static int require_srbd(srbd_ctx* ctx) {
  if (!ctx) return SRBD_ERR_ARG;
  if (!ctx->is_srbd) return fail(ctx, SRBD_ERR_ARG, "NMPC-level calls need nx=nu=12, ng=24, no box constraints");
  return SRBD_OK;
}

int srbd_upload_traj(srbd_ctx* ctx, const double* x, const double* u, const double* xref, const double* x0,
                     const uint8_t* contact) {
  if (int rc = require_srbd(ctx)) return rc;
  if (!x || !u || !xref || !x0) return fail(ctx, SRBD_ERR_ARG, "null trajectory pointer");
  const size_t B = ctx->B, S = ctx->L.N + 1, N = ctx->L.N, D = sizeof(double);
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(ctx->d_x, x, B * S * 12 * D, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_u, u, B * N * 12 * D, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_xref, xref, B * S * 12 * D, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_x0abs, x0, B * 12 * D, cudaMemcpyHostToDevice, ctx->stream));
  ctx->have_contact = contact != nullptr;
  if (contact) CU(cudaMemcpyAsync(ctx->d_contact, contact, B * N * 2, cudaMemcpyHostToDevice, ctx->stream));
  ctx->traj_version++;
  return SRBD_OK;
}

int srbd_download_traj(srbd_ctx* ctx, double* x, double* u) {
  if (int rc = require_srbd(ctx)) return rc;
  const size_t B = ctx->B, S = ctx->L.N + 1, N = ctx->L.N, D = sizeof(double);
  CU(cudaSetDevice(ctx->device));
  if (x) CU(cudaMemcpyAsync(x, ctx->d_x, B * S * 12 * D, cudaMemcpyDeviceToHost, ctx->stream));
  if (u) CU(cudaMemcpyAsync(u, ctx->d_u, B * N * 12 * D, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

// K1.  raw0: also the raw stage-0 blocks (Riccati exports / getters).  dense: also the dense BAbt records -- the throughput
// instantiation of the SRBD K3 variant streams the dyn records and takes the model constants from d_gconst, so the dense
// records (2.7 KB per stage) are written only when something will read them: the generic kernel, the small-batch / exporting
// instantiations of the variant, a getter (ensure_babt), or -- for the listed QPs only -- the rescue pass (qlist).
static int launch_linearize(srbd_ctx* ctx, bool raw0, bool dense, const int* qlist = nullptr, const int* qcount = nullptr) {
  CU(cudaSetDevice(ctx->device));
  LinParams p{};
  p.B = ctx->B; p.N = ctx->L.N;
  p.x = ctx->d_x; p.u = ctx->d_u; p.x0 = ctx->d_x0abs;
  p.babt = ctx->d_babt; p.defect = ctx->d_defect; p.raw0 = raw0 ? ctx->d_raw0 : nullptr; p.dx0 = ctx->d_x0;
  p.gdyn = ctx->d_gdyn;
  p.gconst = qlist ? nullptr : ctx->d_gconst;
  p.dense = dense ? 1 : 0;
  p.qlist = qlist; p.qcount = qcount;
  p.run_gate = ctx->cur_gate;
  const long long total = (long long)p.B * p.N;
  const int grid = (int)((total + kLinThreads - 1) / kLinThreads);   // (work list: the blocks beyond it return at once)
  linearize_kernel<<<grid, kLinThreads, 0, ctx->stream>>>(p, ctx->d_model);
  ctx->launches++;
  CU(cudaGetLastError());
  if (!qlist) {
    ctx->raw0_valid = raw0;
    ctx->gdyn_valid = ctx->d_gdyn != nullptr;
    ctx->babt_dense_valid = dense;
    ctx->lin_traj_version = ctx->traj_version;
  }
  return SRBD_OK;
}

// does anything read the dense BAbt records of this context's linearizations?  (SRBD_K1_DENSE=1: always write them)
static bool want_babt_dense(const srbd_ctx* ctx) {
  const char* d = std::getenv("SRBD_K1_DENSE");
  const char* cge = std::getenv("SRBD_K3_CG");
  return (d && d[0] == '1') || (cge && cge[0] == '0') || want_dense(ctx) || ctx->B < ctx->sm_count || ctx->L.N < 2;
}

// the dense BAbt records for whoever reads them after a dyn-only linearization: re-runs K1 on the unchanged trajectory
static int ensure_babt(srbd_ctx* ctx) {
  if (!ctx->gdyn_valid || ctx->babt_dense_valid) return SRBD_OK;   // (uploads: pack_kernel wrote dense records)
  if (ctx->lin_traj_version != ctx->traj_version)
    return fail(ctx, SRBD_ERR_STATE, "the dense BAbt records were not written by srbd_linearize (throughput path) and the "
                                     "trajectory has changed since (line search / upload): read them before, or set "
                                     "SRBD_K1_DENSE=1");
  return launch_linearize(ctx, ctx->raw0_valid, true);
}

int srbd_linearize(srbd_ctx* ctx) {
  if (int rc = require_srbd(ctx)) return rc;
  return launch_linearize(ctx, want_dense(ctx), want_babt_dense(ctx));
}

// K2.  dense: also RSQrq / DCt / d / dmask / raw stage-0 blocks; qlist / qcount (device): only those QPs
static int launch_assemble(srbd_ctx* ctx, int mode, bool dense, const int* qlist, const int* qcount) {
  CU(cudaSetDevice(ctx->device));
  AsmParams p{};
  p.B = ctx->B; p.N = ctx->L.N; p.mode = mode;
  p.x = ctx->d_x; p.u = ctx->d_u; p.xref = ctx->d_xref; p.contact = ctx->have_contact ? ctx->d_contact : nullptr;
  p.rsq = ctx->d_rsq; p.srec = ctx->d_srec; p.dct = ctx->d_dct; p.d = ctx->d_d; p.dmask = ctx->d_dmask; p.raw0 = ctx->d_raw0;
  p.fcon = nullptr;
  p.qlist = qlist; p.qcount = qcount;
  p.run_gate = ctx->cur_gate;
  const long long total = (long long)p.B * (p.N + 1);
  int grid = (int)((total + kAsmThreads - 1) / kAsmThreads);
  if (qlist && grid > 2 * ctx->sm_count) grid = 2 * ctx->sm_count;  // a rescue list is short (grid-stride loop inside)
  if (mode == SRBD_HARD_INEQ) {
    if (dense) assemble_kernel<true, SRBD_HARD_INEQ><<<grid, kAsmThreads, 0, ctx->stream>>>(p, ctx->d_model);
    else assemble_kernel<false, SRBD_HARD_INEQ><<<grid, kAsmThreads, 0, ctx->stream>>>(p, ctx->d_model);
  } else {
    if (dense) assemble_kernel<true, SRBD_BARRIER_SOFT><<<grid, kAsmThreads, 0, ctx->stream>>>(p, ctx->d_model);
    else assemble_kernel<false, SRBD_BARRIER_SOFT><<<grid, kAsmThreads, 0, ctx->stream>>>(p, ctx->d_model);
  }
  ctx->launches++;
  CU(cudaGetLastError());
  return SRBD_OK;
}

int srbd_assemble(srbd_ctx* ctx, int mode) {
  if (int rc = require_srbd(ctx)) return rc;
  if (mode != SRBD_BARRIER_SOFT && mode != SRBD_HARD_INEQ) return fail(ctx, SRBD_ERR_ARG, "bad assemble mode");
  const bool dense = want_dense(ctx);
  if (int rc = launch_assemble(ctx, mode, dense, nullptr, nullptr)) return rc;
  ctx->packed = true;
  ctx->have_init = false;
  ctx->assembled_mode = mode;
  ctx->dense_valid = dense;
  ctx->asm_traj_version = ctx->traj_version;
  return SRBD_OK;
}

// The dense records of the assembled QP (and the raw stage-0 blocks), for whoever reads them after a compact-only
// assembly: re-runs K1 / K2 on the trajectory the QP was assembled from.
static int ensure_dense(srbd_ctx* ctx) {
  if (ctx->assembled_mode < 0) return SRBD_OK;            // srbd_qp_upload data: pack_kernel always writes dense records
  if (ctx->dense_valid && ctx->raw0_valid) return SRBD_OK;
  if (ctx->asm_traj_version != ctx->traj_version)
    return fail(ctx, SRBD_ERR_STATE, "the dense QP records were not written by srbd_assemble (throughput path) and the "
                                     "trajectory has changed since (line search / upload): read them before, or set "
                                     "SRBD_K2_DENSE=1");
  if (!ctx->raw0_valid)
    if (int rc = launch_linearize(ctx, true, true)) return rc;
  if (!ctx->dense_valid) {
    if (int rc = launch_assemble(ctx, ctx->assembled_mode, true, nullptr, nullptr)) return rc;
    ctx->dense_valid = true;
  }
  return SRBD_OK;
}

int srbd_download_linearization(srbd_ctx* ctx, double* A, double* Bm, double* b, double* defect) {
  if (int rc = require_srbd(ctx)) return rc;
  if (!ctx->raw0_valid && (A || b)) {   // stage 0's A and un-embedded b come from the raw blocks
    if (ctx->asm_traj_version != ctx->traj_version && ctx->assembled_mode >= 0)
      return fail(ctx, SRBD_ERR_STATE, "raw stage-0 blocks were not written (throughput path) and the trajectory has changed");
    if (int rc = launch_linearize(ctx, true, true)) return rc;
  }
  if (int rc = ensure_babt(ctx)) return rc;
  const QpLayout& L = ctx->L;
  const size_t B = ctx->B, N = L.N;
  std::vector<double> h(B * N * L.babt_stride);
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(h.data(), ctx->d_babt, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (defect) CU(cudaMemcpyAsync(defect, ctx->d_defect, B * N * 12 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (size_t q = 0; q < B; ++q)
    for (size_t k = 0; k < N; ++k) {
      const double* r = h.data() + (q * N + k) * L.babt_stride;
      const int nuk = 12, nxk = k > 0 ? 12 : 0, n = nuk + nxk;
      for (int j = 0; j < 12; ++j) {
        for (int i = 0; i < 12; ++i) {
          if (Bm) Bm[(q * N + k) * 144 + j + 12 * i] = r[pm_index(i, j, L.babt_cn)];
          // stage 0 carries no A^T rows (nx[0] := 0): report A0 from the raw block instead (below)
          if (A && k > 0) A[(q * N + k) * 144 + j + 12 * i] = r[pm_index(nuk + i, j, L.babt_cn)];
        }
        if (b) b[(q * N + k) * 12 + j] = r[pm_index(n, j, L.babt_cn)];
      }
    }
  if (A || b) {  // stage 0: raw A0 and the un-embedded b0
    std::vector<double> raw(B * kRaw0Stride);
    CU(cudaMemcpy(raw.data(), ctx->d_raw0, raw.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t q = 0; q < B; ++q) {
      if (A) std::memcpy(A + q * N * 144, raw.data() + q * kRaw0Stride, 144 * sizeof(double));
      if (b) std::memcpy(b + q * N * 12, raw.data() + q * kRaw0Stride + 288, 12 * sizeof(double));
    }
  }
  return SRBD_OK;
}

int srbd_download_qp(srbd_ctx* ctx, double* Q, double* S, double* R, double* q, double* r, double* D, double* lg,
                     double* lg_mask) {
  if (int rc = require_srbd(ctx)) return rc;
  if (!ctx->packed) return fail(ctx, SRBD_ERR_STATE, "assemble first");
  if (int rc = ensure_dense(ctx)) return rc;
  const QpLayout& L = ctx->L;
  const size_t B = ctx->B, N = L.N, S1 = N + 1;
  std::vector<double> hr(B * S1 * L.rsq_stride), hd(B * S1 * L.dct_stride), hv(B * S1 * L.d_stride), hm(B * S1 * L.d_stride);
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(hr.data(), ctx->d_rsq, hr.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hd.data(), ctx->d_dct, hd.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hv.data(), ctx->d_d, hv.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hm.data(), ctx->d_dmask, hm.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (size_t b = 0; b < B; ++b)
    for (size_t k = 0; k <= N; ++k) {
      const double* rs = hr.data() + (b * S1 + k) * L.rsq_stride;
      const int nuk = k < N ? 12 : 0, nxk = k > 0 ? 12 : 0, n = nuk + nxk;
      for (int j = 0; j < 12; ++j) {
        for (int i = 0; i < 12; ++i) {
          if (k < N && R) R[(b * N + k) * 144 + i + 12 * j] = rs[pm_index(i >= j ? i : j, i >= j ? j : i, L.rsq_cn)];
          if (Q) {  // stage 0 has no state block in the packed record: Q0 does not enter the embedded QP
            double v = 0.0;
            if (nxk) v = rs[pm_index(nuk + (i >= j ? i : j), nuk + (i >= j ? j : i), L.rsq_cn)];
            else v = (i == j) ? ctx->model.Q[i] : 0.0;
            Q[(b * S1 + k) * 144 + i + 12 * j] = v;
          }
          if (k < N && S) S[(b * N + k) * 144 + i + 12 * j] = (nxk ? rs[pm_index(nuk + j, i, L.rsq_cn)] : 0.0);
        }
        if (k < N && r) r[(b * N + k) * 12 + j] = rs[pm_index(n, j, L.rsq_cn)];
        if (q && nxk) q[(b * S1 + k) * 12 + j] = rs[pm_index(n, nuk + j, L.rsq_cn)];
      }
      if (k < N) {
        const double* dc = hd.data() + (b * S1 + k) * L.dct_stride;
        for (int g = 0; g < 24; ++g) {
          for (int i = 0; i < 12; ++i)
            if (D) D[(b * N + k) * 288 + g + 24 * i] = dc[pm_index(i, g, L.dct_cn)];
          if (lg) lg[(b * N + k) * 24 + g] = hv[(b * S1 + k) * L.d_stride + g];
          if (lg_mask) lg_mask[(b * N + k) * 24 + g] = hm[(b * S1 + k) * L.d_stride + g];
        }
      }
    }
  if (q) {  // q0 from the raw block
    std::vector<double> raw(B * kRaw0Stride);
    CU(cudaMemcpy(raw.data(), ctx->d_raw0, raw.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t b = 0; b < B; ++b) std::memcpy(q + b * S1 * 12, raw.data() + b * kRaw0Stride + 588, 12 * sizeof(double));
  }
  return SRBD_OK;
}

// ---- QP level -------------------------------------------------------------------------------------
int srbd_qp_upload_layout(srbd_ctx* ctx, int d_shared) {
  if (!ctx) return SRBD_ERR_ARG;
  if (ctx->upload_d_shared != (d_shared != 0) && ctx->raw_alloc) {   // per-field device buffers were sized for the other layout
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    for (void*& q : ctx->raw_dev) { if (q) cudaFree(q); q = nullptr; }
    ctx->raw_alloc = false;
  }
  ctx->upload_d_shared = d_shared != 0;
  return SRBD_OK;
}

int srbd_qp_upload(srbd_ctx* ctx, const srbd_qp_host* qp) {
  if (!ctx || !qp) return SRBD_ERR_ARG;
  const QpLayout& L0 = ctx->L;
  const srbd_qp_dims& d = ctx->dims;
  if (!qp->A || !qp->Bm || !qp->b || !qp->Q || !qp->R || !qp->q || !qp->r || !qp->x0)
    return fail(ctx, SRBD_ERR_ARG, "A, B, b, Q, R, q, r and x0 are required");
  if (d.nbx > 0 && (!qp->idxbx || !qp->lbx || !qp->ubx)) return fail(ctx, SRBD_ERR_ARG, "nbx > 0 needs idxbx, lbx, ubx");
  if (d.nbu > 0 && (!qp->idxbu || !qp->lbu || !qp->ubu)) return fail(ctx, SRBD_ERR_ARG, "nbu > 0 needs idxbu, lbu, ubu");
  if (d.ng > 0 && (!qp->D || !qp->lg || !qp->ug)) return fail(ctx, SRBD_ERR_ARG, "ng > 0 needs D, lg, ug");
  if (d.ngN > 0 && (!qp->CN || !qp->lgN || !qp->ugN)) return fail(ctx, SRBD_ERR_ARG, "ngN > 0 needs CN, lgN, ugN");
  QpLayout L;
  if (make_layout(d, qp->idxbx, qp->idxbu, &L) != 0) return fail(ctx, SRBD_ERR_ARG, "bad idxbx / idxbu");
  ctx->L = L;
  (void)L0;
  CU(cudaSetDevice(ctx->device));
  const size_t B = ctx->B, N = L.N, S = N + 1, nx = L.nx, nu = L.nu;
  struct F { const double* srbd_qp_host::*m; size_t n; };
  const F fields[] = {
      {&srbd_qp_host::A, B * N * nx * nx}, {&srbd_qp_host::Bm, B * N * nx * nu}, {&srbd_qp_host::b, B * N * nx},
      {&srbd_qp_host::Q, B * S * nx * nx}, {&srbd_qp_host::S, B * N * nu * nx}, {&srbd_qp_host::R, B * N * nu * nu},
      {&srbd_qp_host::q, B * S * nx}, {&srbd_qp_host::r, B * N * nu},
      {&srbd_qp_host::lbx, B * S * L.nbx}, {&srbd_qp_host::ubx, B * S * L.nbx},
      {&srbd_qp_host::lbx_mask, B * S * L.nbx}, {&srbd_qp_host::ubx_mask, B * S * L.nbx},
      {&srbd_qp_host::lbu, B * N * L.nbu}, {&srbd_qp_host::ubu, B * N * L.nbu},
      {&srbd_qp_host::lbu_mask, B * N * L.nbu}, {&srbd_qp_host::ubu_mask, B * N * L.nbu},
      {&srbd_qp_host::C, B * N * L.ng * nx}, {&srbd_qp_host::D, ctx->upload_d_shared ? (size_t)L.ng * nu : B * N * L.ng * nu},
      {&srbd_qp_host::lg, B * N * L.ng}, {&srbd_qp_host::ug, B * N * L.ng},
      {&srbd_qp_host::lg_mask, B * N * L.ng}, {&srbd_qp_host::ug_mask, B * N * L.ng},
      {&srbd_qp_host::CN, B * L.ngN * nx}, {&srbd_qp_host::lgN, B * L.ngN}, {&srbd_qp_host::ugN, B * L.ngN},
      {&srbd_qp_host::lgN_mask, B * L.ngN}, {&srbd_qp_host::ugN_mask, B * L.ngN},
      {&srbd_qp_host::x0, B * nx}};
  const int nf = sizeof(fields) / sizeof(fields[0]);
  // Host fields that lie in ONE contiguous range (the facades flatten into a pinned arena) travel in a single H2D copy
  const char *lo = nullptr, *hi = nullptr;
  size_t sum = 0;
  auto span = [&](const double* ptr, size_t n) {
    if (!ptr || !n) return;
    const char* a = reinterpret_cast<const char*>(ptr);
    if (!lo || a < lo) lo = a;
    if (!hi || a + n * sizeof(double) > hi) hi = a + n * sizeof(double);
    sum += n * sizeof(double);
  };
  for (int i = 0; i < nf; ++i) span(qp->*(fields[i].m), fields[i].n);
  span(qp->x_init, qp->x_init && qp->u_init ? B * S * nx : 0);
  span(qp->u_init, qp->x_init && qp->u_init ? B * N * nu : 0);
  const bool contiguous = lo && (size_t)(hi - lo) <= sum + 64 * (size_t)(nf + 2);
  srbd_qp_host dq{};
  auto dev_of = [&](const double* src) { return reinterpret_cast<const double*>(ctx->d_in + (reinterpret_cast<const char*>(src) - lo)); };
  if (contiguous) {
    const size_t bytes = (size_t)(hi - lo);
    if (bytes > ctx->in_bytes) {
      if (ctx->d_in) cudaFree(ctx->d_in);
      ctx->d_in = nullptr; ctx->in_bytes = 0;
      CU(cudaMalloc(reinterpret_cast<void**>(&ctx->d_in), bytes));
      ctx->in_bytes = bytes;
    }
    CU(cudaMemcpyAsync(ctx->d_in, lo, bytes, cudaMemcpyHostToDevice, ctx->stream));
    for (int i = 0; i < nf; ++i) {
      const double* src = qp->*(fields[i].m);
      if (src && fields[i].n) dq.*(fields[i].m) = dev_of(src);
    }
    CU(cudaMemcpyAsync(ctx->d_x0, dev_of(qp->x0), B * nx * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    if (!ctx->raw_alloc) {
      ctx->raw_dev.assign(nf, nullptr);
      for (int i = 0; i < nf; ++i)
        if (fields[i].n) CU(cudaMalloc(&ctx->raw_dev[i], fields[i].n * sizeof(double)));
      ctx->raw_alloc = true;
    }
    for (int i = 0; i < nf; ++i) {
      const double* src = qp->*(fields[i].m);
      if (src && fields[i].n) {
        CU(cudaMemcpyAsync(ctx->raw_dev[i], src, fields[i].n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        dq.*(fields[i].m) = (const double*)ctx->raw_dev[i];
      }
    }
    CU(cudaMemcpyAsync(ctx->d_x0, qp->x0, B * nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->have_init = false;
  if (qp->x_init && qp->u_init) {
    const cudaMemcpyKind kd = contiguous ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CU(cudaMemcpyAsync(ctx->d_xinit, contiguous ? dev_of(qp->x_init) : qp->x_init, B * S * nx * sizeof(double), kd, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_uinit, contiguous ? dev_of(qp->u_init) : qp->u_init, B * N * nu * sizeof(double), kd, ctx->stream));
    ctx->have_init = true;
  }
  PackParams p{};
  p.L = L; p.B = ctx->B; p.qp = dq;
  p.babt = ctx->d_babt; p.rsq = ctx->d_rsq; p.dct = ctx->d_dct; p.d = ctx->d_d; p.dmask = ctx->d_dmask;
  p.raw0 = ctx->d_raw0; p.raw0_stride = raw0_stride(L); p.r0raw = ctx->d_r0raw;
  p.d_stride = ctx->upload_d_shared ? 0 : L.ng * L.nu;
  const long long total = (long long)ctx->B * (L.N + 1);
  pack_kernel<<<(int)((total + 3) / 4), 128, 0, ctx->stream>>>(p);
  ctx->launches++;
  CU(cudaGetLastError());
  ctx->packed = true;
  ctx->assembled_mode = -1;
  ctx->gdyn_valid = false;
  ctx->upload_variant_ok = false;
  if (ctx->is_srbd) {  // SRBD dimensions: does the data have K2's structure?  (decided on the device, aux_kernels.cuh)
    DetectParams dp{};
    dp.B = ctx->B; dp.N = L.N; dp.qp = dq; dp.srec = ctx->d_srec; dp.model = ctx->d_model_qp; dp.bad = ctx->d_flag;
    dp.d_stride = ctx->upload_d_shared ? 0 : 288;
    CU(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    detect_srbd_kernel<<<(int)((total + 3) / 4), 128, 0, ctx->stream>>>(dp);
    ctx->launches++;
    CU(cudaGetLastError());
    ctx->upload_variant_ok = true;
  }
  return SRBD_OK;
}

// K3, generic kernel (any hpipm::OcpQp data).  qlist / qcount (device): solve only those QPs and keep the batch
// statistics accumulated so far (the rescue pass of the SRBD variant); null: the whole batch.
static int launch_generic(srbd_ctx* ctx, const int* qlist, const int* qcount, const int* gate = nullptr, int gate_value = 0,
                          bool keep_stats = false) {
  const QpLayout& L = ctx->L;
  const size_t B = ctx->B, S = L.N + 1, N = L.N;
  if (ctx->export_ric && !ctx->d_P) {
    CU(dalloc(&ctx->d_P, B * S * L.nx * L.nx)); CU(dalloc(&ctx->d_p, B * S * L.nx));
    CU(dalloc(&ctx->d_K, B * N * L.nu * L.nx)); CU(dalloc(&ctx->d_k, B * N * L.nu));
      CU(dalloc(&ctx->d_Lr0, B * L.nu * L.nu));
  }
  if (ctx->export_stat && !ctx->d_stat) CU(dalloc(&ctx->d_stat, B * (size_t)ctx->stat_rows * SRBD_STAT_M));
  if (!qlist)
    if (int rc = ensure_babt(ctx)) return rc;   // (a work list: the caller has linearized the listed QPs densely)
  IpmParams p{};
  p.L = L; p.a = ctx->args; p.B = ctx->B;
  p.babt = ctx->d_babt; p.rsq = ctx->d_rsq; p.dct = ctx->d_dct; p.d = ctx->d_d; p.dmask = ctx->d_dmask;
  p.x_init = ctx->have_init ? ctx->d_xinit : nullptr;
  p.u_init = ctx->have_init ? ctx->d_uinit : nullptr;
  if (ctx->warm_from_solution) {  // every QP reads its guess before it writes its solution: in place
    p.x_init = ctx->d_sol_x;
    p.u_init = ctx->d_sol_u;
  }
  p.x0 = ctx->d_x0; p.raw0 = ctx->d_raw0; p.ws = ctx->d_ws; p.counter = ctx->d_counter;
  p.sol_x = ctx->d_sol_x; p.sol_u = ctx->d_sol_u; p.sol_pi = ctx->d_sol_pi; p.sol_lam = ctx->d_sol_lam; p.sol_t = ctx->d_sol_t;
  if (ctx->export_ric) { p.ric_P = ctx->d_P; p.ric_p = ctx->d_p; p.ric_K = ctx->d_K; p.ric_k = ctx->d_k; p.ric_Lr0 = ctx->d_Lr0; }
  p.iter = ctx->d_iter; p.status = ctx->d_status; p.res_max = ctx->d_resmax;
  p.stat = ctx->export_stat ? ctx->d_stat : nullptr;
  p.stat_rows = ctx->stat_rows;
  p.bstats = ctx->d_bstats;
  CU(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
  if (!qlist && !keep_stats) CU(cudaMemsetAsync(ctx->d_bstats, 0, sizeof(srbd_batch_stats), ctx->stream));
  p.qlist = qlist; p.qcount = qcount;
  p.gate = gate; p.gate_value = gate_value;
  p.run_gate = ctx->cur_gate; p.frozen = ctx->sqp_loop ? ctx->d_conv : nullptr;
  KernelChoice kc = pick_kernel(L);
  const int grid = qlist ? (ctx->grid < ctx->sm_count ? ctx->grid : ctx->sm_count) : ctx->grid;  // a rescue list is short
  // fewer CTAs than SMs: the workspace of each CTA's QP in its dynamic shared memory, when it fits (IpmParams::ws_in_smem)
  const bool refine = ctx->args.itref_pred_max > 0 || ctx->args.itref_corr_max > 0;
  const size_t ws_bytes = (size_t)(refine ? L.ws_size : L.ws_size_core) * sizeof(double);
  const char* nosm = std::getenv("SRBD_K3_WS_GLOBAL");
  size_t dyn = 0;
  if (grid <= ctx->sm_count && ws_bytes + 24 * 1024 <= (size_t)ctx->smem_optin && !(nosm && nosm[0] == '1')) {
    if (cudaFuncSetAttribute(kc.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws_bytes) == cudaSuccess) dyn = ws_bytes;
    else cudaGetLastError();
  }
  p.ws_in_smem = dyn ? 1 : 0;
  kc.fn<<<grid, 32, dyn, ctx->stream>>>(p);
  ctx->launches++;
  CU(cudaGetLastError());
  ctx->solved = true;
  if (!qlist) { ctx->ric_valid = ctx->export_ric; ctx->stat_valid = ctx->export_stat; }
  return SRBD_OK;
}

// K3 for QPs assembled by K2: the SRBD throughput variant (same algorithm as the generic kernel)
static int solve_srbd_variant(srbd_ctx* ctx, const ModelDev* model, const int* gate = nullptr) {
  const QpLayout& L = ctx->L;
  constexpr int kTeamSmem = v2::kSmemBytes + kTeamShared * 8;
  // LATENCY mode: with fewer QPs than SMs every QP gets a whole CTA (the team instantiation: the residual sweep split over
  // the warps, bit-identical results).  SRBD_K3_TEAM=0: one warp per QP as in the throughput mode.
  const char* te = std::getenv("SRBD_K3_TEAM");
  const bool team = ctx->B < ctx->sm_count && !(te && te[0] == '0');
  if (!ctx->d_ws2) {
    CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes));
    CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA, 0, 0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes));
    CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes));
    CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes));
    CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 0, v2::kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTeamSmem));
    CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 1, v2::kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTeamSmem));
    CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 0, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes));
    CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 0, v2::kWarps, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTeamSmem));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ipm_srbd_kernel<SRBD_K3_TMA, 0>, 32 * v2::kWarps, v2::kSmemBytes));
    if (occ < 1) return fail(ctx, SRBD_ERR_CUDA, "ipm_srbd_kernel does not fit on this device");
    if (const char* cap = std::getenv("SRBD_K3_CTAS_PER_SM")) {  // tuning knob: fewer resident QPs = higher L2 hit rate
      const int c = std::atoi(cap);
      if (c >= 1 && c < occ) occ = c;
    }
    long long g = (long long)occ * ctx->sm_count;
    if (const char* gs = std::getenv("SRBD_K3_GRID")) {  // tuning knob: number of resident CTAs (4 QPs each)
      const long long c = std::atoll(gs);
      if (c >= 1 && c < g) g = c;
    }
    // Sparse batches (at least as many QPs as SMs, fewer than resident warps; BASELINE config 2: 1024 QPs): the full grid of
    // the kSpread instantiation, which keeps exactly B warps spread evenly over the SMs (6-8 solves side by side on every
    // SM instead of 12 on some and 6 on the others: 1024 QPs 3.00 -> 2.55 ms).  Everything else: no more CTAs than needed.
    const long long need = (ctx->B + v2::kWarps - 1) / v2::kWarps;
    const char* sp = std::getenv("SRBD_K3_SPREAD");   // =0: CTAs filled one after the other as for full batches (A/B runs)
    ctx->spread = g > need && ctx->B >= ctx->sm_count && !(sp && sp[0] == '0');
    if (g > need && !ctx->spread) g = need;
    ctx->grid2 = (int)g;
    if (ctx->spread) {
      CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA, 0, 0, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes));
      CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA, 0, 0, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes));
      CU(cudaFuncSetAttribute(ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 0, 0, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes));
    }
    CU(dalloc(&ctx->d_ws2, (size_t)ctx->grid2 * v2::kWarps * (size_t)(L.N + 1) * (v2::kStage + v2::kAlt)));
    CU(dalloc(&ctx->d_retry, (size_t)ctx->B + 1));
    CU(dalloc(&ctx->d_retry2, (size_t)ctx->B + 1));
  }
  // Rescue pass: a QP on which the variant runs to iter_max (about 1 in 1e5 sits on the rounding floor of the blocked
  // triangular solves) is solved again by the generic kernel, whose row-by-row substitution is as
  // accurate as the reference's trsv.  Only iter_max: the min-step / NaN endings seen on the long-horizon workload
  // (config 4, N = 100) end the same way in the generic kernel, at 50 ms per one-warp solve.  SRBD_K3_NO_RESCUE=1
  // switches the pass off (diagnosis).
  const char* nr = std::getenv("SRBD_K3_NO_RESCUE");
  const bool rescue = !(nr && nr[0] == '1');
  SrbdIpmParams p{};
  p.B = ctx->B; p.N = L.N; p.a = ctx->args;
  p.babt = ctx->d_babt; p.srec = ctx->d_srec; p.x0 = ctx->d_x0; p.gdyn = ctx->d_gdyn; p.gconst = ctx->d_gconst; p.asm_mode = ctx->assembled_mode;
  p.model = model; p.ws = ctx->d_ws2;
  p.gate = gate; p.gate_value = 0;
  p.run_gate = ctx->cur_gate; p.frozen = ctx->sqp_loop ? ctx->d_conv : nullptr; p.ws_size = (L.N + 1) * (v2::kStage + v2::kAlt); p.counter = ctx->d_counter;
  p.sol_x = ctx->d_sol_x; p.sol_u = ctx->d_sol_u; p.sol_pi = ctx->d_sol_pi; p.sol_lam = ctx->d_sol_lam; p.sol_t = ctx->d_sol_t;
  p.iter = ctx->d_iter; p.status = ctx->d_status; p.res_max = ctx->d_resmax; p.bstats = ctx->d_bstats;
  if (rescue) {
    p.retry_list = ctx->d_retry; p.retry_count = ctx->d_retry + ctx->B;
    CU(cudaMemsetAsync(ctx->d_retry + ctx->B, 0, sizeof(int), ctx->stream));
  }
  if (ctx->export_ric) {   // (QP-level uploads only: settings_allow_variant)
    const size_t B = ctx->B, S = L.N + 1, N = L.N;
    if (!ctx->d_P) {
      CU(dalloc(&ctx->d_P, B * S * L.nx * L.nx)); CU(dalloc(&ctx->d_p, B * S * L.nx));
      CU(dalloc(&ctx->d_K, B * N * L.nu * L.nx)); CU(dalloc(&ctx->d_k, B * N * L.nu));
      CU(dalloc(&ctx->d_Lr0, B * L.nu * L.nu));
    }
    p.ric_P = ctx->d_P; p.ric_p = ctx->d_p; p.ric_K = ctx->d_K; p.ric_k = ctx->d_k; p.ric_Lr0 = ctx->d_Lr0;
    p.raw0 = ctx->d_raw0; p.raw0_stride = raw0_stride(L);
  }
  if (ctx->export_stat) {
    if (!ctx->d_stat) CU(dalloc(&ctx->d_stat, (size_t)ctx->B * ctx->stat_rows * SRBD_STAT_M));
    p.stat = ctx->d_stat; p.stat_rows = ctx->stat_rows;
  }
  CU(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_bstats, 0, sizeof(srbd_batch_stats), ctx->stream));
  // batches of fewer QPs than SMs have nothing to hide a bulk copy's latency behind: the cp.async-only instantiation
  // with the facade's exports: the exporting instantiations (cp.async tile engine); their rescue list goes straight to the
  // generic kernel, which exports too
  const bool exports = ctx->export_ric || ctx->export_stat;
  // QPs linearized by K1 (not QP-level uploads, whose A / B are arbitrary): the BAbt tiles keep the model constants
  // resident and stream K1's dyn records only (ipm_srbd.cuh: kCG).  SRBD_K3_CG=0: dense records (A/B runs).
  const char* cge = std::getenv("SRBD_K3_CG");
  const bool compact = ctx->assembled_mode >= 0 && ctx->gdyn_valid && L.N >= 2 && !(cge && cge[0] == '0');
  if (!compact)
    if (int rc = ensure_babt(ctx)) return rc;   // every other instantiation streams the dense BAbt records
  if (exports && team) ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 0, v2::kWarps, true><<<ctx->B, 32 * v2::kWarps, kTeamSmem, ctx->stream>>>(p);
  else if (exports && ctx->spread) ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 0, 0, true, false, true><<<ctx->grid2, 32 * v2::kWarps, v2::kSmemBytes, ctx->stream>>>(p);
  else if (exports) ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 0, 0, true><<<ctx->grid2, 32 * v2::kWarps, v2::kSmemBytes, ctx->stream>>>(p);
  else if (team) ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 0, v2::kWarps><<<ctx->B, 32 * v2::kWarps, kTeamSmem, ctx->stream>>>(p);
  else if (ctx->B < ctx->sm_count) ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 0><<<ctx->grid2, 32 * v2::kWarps, v2::kSmemBytes, ctx->stream>>>(p);
  else if (compact && ctx->spread) ipm_srbd_kernel<SRBD_K3_TMA, 0, 0, false, true, true><<<ctx->grid2, 32 * v2::kWarps, v2::kSmemBytes, ctx->stream>>>(p);
  else if (ctx->spread) ipm_srbd_kernel<SRBD_K3_TMA, 0, 0, false, false, true><<<ctx->grid2, 32 * v2::kWarps, v2::kSmemBytes, ctx->stream>>>(p);
  else if (compact) ipm_srbd_kernel<SRBD_K3_TMA, 0, 0, false, true><<<ctx->grid2, 32 * v2::kWarps, v2::kSmemBytes, ctx->stream>>>(p);
  else ipm_srbd_kernel<SRBD_K3_TMA, 0><<<ctx->grid2, 32 * v2::kWarps, v2::kSmemBytes, ctx->stream>>>(p);
  ctx->launches++;
  CU(cudaGetLastError());
  ctx->solved = true;
  ctx->ric_valid = ctx->export_ric;   // (P, p, K, k, pi[0]: only on request; see settings_allow_variant)
  ctx->stat_valid = ctx->export_stat;
  if (rescue && ctx->gdyn_valid && !ctx->babt_dense_valid)   // the rescue kernels read dense BAbt records: the listed QPs only
    if (int rc = launch_linearize(ctx, false, true, ctx->d_retry, ctx->d_retry + ctx->B)) return rc;
  if (rescue && exports) {
    if (ctx->assembled_mode >= 0 && !ctx->dense_valid)
      if (int rc = launch_assemble(ctx, ctx->assembled_mode, true, ctx->d_retry, ctx->d_retry + ctx->B)) return rc;
    return launch_generic(ctx, ctx->d_retry, ctx->d_retry + ctx->B);
  }
  if (rescue) {
    // Stage 1: the listed QPs again in the SAME tensor-core kernel with the other rounding of the inverse pivots (kPivot =
    // 1): a QP that runs to iter_max sits on a knife edge of the rounding floor, and each rounding has its own (about 3
    // per million, scripts/count_iter_max.py).  An empty kernel unless the list is non-empty; 1.4 ms for one QP.
    SrbdIpmParams r = p;
    r.qlist = ctx->d_retry; r.qcount = ctx->d_retry + ctx->B;
    r.retry_list = ctx->d_retry2; r.retry_count = ctx->d_retry2 + ctx->B;
    CU(cudaMemsetAsync(ctx->d_retry2 + ctx->B, 0, sizeof(int), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), ctx->stream));
    const int g1 = ctx->grid2 < ctx->sm_count ? ctx->grid2 : ctx->sm_count;
    if (team) ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 1, v2::kWarps><<<ctx->B, 32 * v2::kWarps, kTeamSmem, ctx->stream>>>(r);
    else ipm_srbd_kernel<SRBD_K3_TMA_SMALL, 1><<<g1, 32 * v2::kWarps, v2::kSmemBytes, ctx->stream>>>(r);
    ctx->launches++;
    CU(cudaGetLastError());
    // Stage 2: what is still on the list goes to the generic kernel (row-by-row substitution, no block inverses), which
    // reads the dense records: write them for the listed QPs only (empty kernels otherwise)
    if (ctx->assembled_mode >= 0 && !ctx->dense_valid)
      if (int rc = launch_assemble(ctx, ctx->assembled_mode, true, ctx->d_retry2, ctx->d_retry2 + ctx->B)) return rc;
    return launch_generic(ctx, ctx->d_retry2, ctx->d_retry2 + ctx->B);
  }
  return SRBD_OK;
}

int srbd_qp_solve(srbd_ctx* ctx) {
  if (!ctx) return SRBD_ERR_ARG;
  if (!ctx->packed) return fail(ctx, SRBD_ERR_STATE, "no QP data: call srbd_qp_upload or srbd_assemble first");
  if (ctx->args.warm_start && !ctx->have_init && !ctx->is_srbd)
    return fail(ctx, SRBD_ERR_ARG, "warm_start=1 needs x_init/u_init (qp_sol[i].x / .u must be pre-sized, "
                                   "hpipm-cpp/src/ocp_qp_ipm_solver.cpp:190-207)");
  CU(cudaSetDevice(ctx->device));
  // QPs assembled by K2 (both modes: BARRIER_SOFT masks every row, which the variant solves as the single unconstrained
  // Riccati pass) take the SRBD tensor-core variant unless a setting needs the generic kernel
  if (ctx->assembled_mode >= 0 && variant_eligible(ctx)) return solve_srbd_variant(ctx, ctx->d_model);
  if (ctx->assembled_mode < 0 && ctx->upload_variant_ok && settings_allow_variant(ctx, true)) {
    // Uploaded QP with the SRBD dimensions: detect_srbd_kernel left *d_flag = 0 if it has K2's structure.  Both kernels
    // are launched, each gated on the flag (no host round trip): the variant (+ its rescue launch) runs if 0, the
    // generic kernel if 1.
    if (int rc = solve_srbd_variant(ctx, ctx->d_model_qp, ctx->d_flag)) return rc;
    return launch_generic(ctx, nullptr, nullptr, ctx->d_flag, 1, true);
  }
  if (int rc = ensure_dense(ctx)) return rc;
  return launch_generic(ctx, nullptr, nullptr);
}

int srbd_download_solution(srbd_ctx* ctx, const srbd_sol_host* sol) {
  if (!ctx || !sol) return SRBD_ERR_ARG;
  if (!ctx->solved) return fail(ctx, SRBD_ERR_STATE, "solve first");
  const QpLayout& L = ctx->L;
  const size_t B = ctx->B, S = L.N + 1, N = L.N, D = sizeof(double);
  CU(cudaSetDevice(ctx->device));
  auto dl = [&](double* dst, const double* src, size_t n) -> cudaError_t {
    return dst ? cudaMemcpyAsync(dst, src, n * D, cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess;
  };
  if ((sol->P || sol->p || sol->K || sol->k) && !(ctx->d_P && ctx->ric_valid))
    return fail(ctx, SRBD_ERR_STATE, "Riccati outputs were not exported by the last solve: call srbd_set_outputs(ctx, 1, ..) "
                                     "before srbd_qp_solve (pi[0] is only produced together with them)");
  CU(dl(sol->x, ctx->d_sol_x, B * S * L.nx)); CU(dl(sol->u, ctx->d_sol_u, B * N * L.nu));
  CU(dl(sol->pi, ctx->d_sol_pi, B * S * L.nx));
  CU(dl(sol->lam, ctx->d_sol_lam, B * (size_t)L.nct)); CU(dl(sol->t, ctx->d_sol_t, B * (size_t)L.nct));
  CU(dl(sol->P, ctx->d_P, B * S * L.nx * L.nx)); CU(dl(sol->p, ctx->d_p, B * S * L.nx));
  CU(dl(sol->K, ctx->d_K, B * N * L.nu * L.nx)); CU(dl(sol->k, ctx->d_k, B * N * L.nu));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

int srbd_download_ric_lr0(srbd_ctx* ctx, double* Lr0) {
  if (!ctx || !Lr0) return SRBD_ERR_ARG;
  if (!ctx->solved) return fail(ctx, SRBD_ERR_STATE, "solve first");
  if (!(ctx->d_Lr0 && ctx->ric_valid))
    return fail(ctx, SRBD_ERR_STATE, "Riccati outputs were not exported by the last solve: call srbd_set_outputs(ctx, 1, ..) first");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(Lr0, ctx->d_Lr0, (size_t)ctx->B * ctx->L.nu * ctx->L.nu * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

int srbd_out_layout(const srbd_ctx* ctx, size_t offs[8], size_t* total_doubles) {
  if (!ctx || !offs || !total_doubles) return SRBD_ERR_ARG;
  for (int i = 0; i < 8; ++i) offs[i] = ctx->out_off[i];
  *total_doubles = ctx->out_total;
  return SRBD_OK;
}

int srbd_download_packed(srbd_ctx* ctx, double* arena, int with_duals) {
  if (!ctx || !arena) return SRBD_ERR_ARG;
  if (!ctx->solved) return fail(ctx, SRBD_ERR_STATE, "solve first");
  CU(cudaSetDevice(ctx->device));
  const size_t n = with_duals ? ctx->out_total : ctx->out_off[6];
  CU(cudaMemcpyAsync(arena, ctx->d_out, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

int srbd_host_alloc(size_t bytes, void** ptr) {
  if (!ptr) return SRBD_ERR_ARG;
  *ptr = nullptr;
  return cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? SRBD_OK : SRBD_ERR_CUDA;
}

int srbd_host_free(void* ptr) {
  if (ptr && cudaFreeHost(ptr) != cudaSuccess) return SRBD_ERR_CUDA;
  return SRBD_OK;
}

int srbd_download_stats(srbd_ctx* ctx, const srbd_stats_host* st) {
  if (!ctx || !st) return SRBD_ERR_ARG;
  if (!ctx->solved) return fail(ctx, SRBD_ERR_STATE, "solve first");
  const size_t B = ctx->B;
  CU(cudaSetDevice(ctx->device));
  if (st->iter) CU(cudaMemcpyAsync(st->iter, ctx->d_iter, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (st->status) CU(cudaMemcpyAsync(st->status, ctx->d_status, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (st->res_max) CU(cudaMemcpyAsync(st->res_max, ctx->d_resmax, B * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (st->stat) {
    if (!ctx->d_stat || !ctx->stat_valid)
      return fail(ctx, SRBD_ERR_STATE, "statistics table was not exported by the last solve: srbd_set_outputs(ctx, .., 1)");
    CU(cudaMemcpyAsync(st->stat, ctx->d_stat, B * (size_t)ctx->stat_rows * SRBD_STAT_M * sizeof(double),
                       cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

int srbd_batch_stats_get(srbd_ctx* ctx, srbd_batch_stats* out) {
  if (!ctx || !out) return SRBD_ERR_ARG;
  if (!ctx->solved) return fail(ctx, SRBD_ERR_STATE, "solve first");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(out, ctx->d_bstats, sizeof(*out), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

// ---- closed-loop batched MPC -----------------------------------------------------------------------
int srbd_mpc_run(srbd_ctx* ctx, const double* A, const double* Bm, const double* b, int plant_shared, const double* x_start,
                 int steps, double* x_traj, double* u_traj, int* iter, int* status) {
  if (!ctx || !A || !Bm || !b || !x_start || steps < 1) return SRBD_ERR_ARG;
  if (!ctx->packed || ctx->assembled_mode >= 0)
    return fail(ctx, SRBD_ERR_STATE, "srbd_mpc_run needs QP data from srbd_qp_upload (the QP stays fixed, x0 changes per step)");
  if (ctx->args.warm_start && !ctx->have_init)
    return fail(ctx, SRBD_ERR_ARG, "warm_start=1 needs x_init/u_init for the first step");
  CU(cudaSetDevice(ctx->device));
  const QpLayout& L = ctx->L;
  const size_t B = ctx->B, nx = L.nx, nu = L.nu, D = sizeof(double);
  const size_t np = plant_shared ? 1 : B;
  if (steps > ctx->mpc_steps_alloc) {
    for (void* q : {(void*)ctx->d_mpc_x, (void*)ctx->d_mpc_u, (void*)ctx->d_mpc_iter, (void*)ctx->d_mpc_status})
      if (q) cudaFree(q);
    ctx->d_mpc_x = ctx->d_mpc_u = nullptr; ctx->d_mpc_iter = ctx->d_mpc_status = nullptr;
    CU(dalloc(&ctx->d_mpc_x, (size_t)(steps + 1) * B * nx)); CU(dalloc(&ctx->d_mpc_u, (size_t)steps * B * nu));
    CU(dalloc(&ctx->d_mpc_iter, (size_t)steps * B)); CU(dalloc(&ctx->d_mpc_status, (size_t)steps * B));
    ctx->mpc_steps_alloc = steps;
  }
  if (!ctx->d_mpc_xcur) CU(dalloc(&ctx->d_mpc_xcur, B * nx));
  if ((int)np > ctx->plant_alloc) {
    for (void* q : {(void*)ctx->d_plantA, (void*)ctx->d_plantB, (void*)ctx->d_plantb})
      if (q) cudaFree(q);
    ctx->d_plantA = ctx->d_plantB = ctx->d_plantb = nullptr;
    CU(dalloc(&ctx->d_plantA, np * nx * nx)); CU(dalloc(&ctx->d_plantB, np * nx * nu)); CU(dalloc(&ctx->d_plantb, np * nx));
    ctx->plant_alloc = (int)np;
  }
  CU(cudaMemcpyAsync(ctx->d_plantA, A, np * nx * nx * D, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_plantB, Bm, np * nx * nu * D, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_plantb, b, np * nx * D, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_mpc_xcur, x_start, B * nx * D, cudaMemcpyHostToDevice, ctx->stream));
  MpcParams p{};
  p.L = L; p.B = ctx->B;
  p.raw0 = ctx->d_raw0; p.raw0_stride = raw0_stride(L); p.r0raw = ctx->d_r0raw;
  p.xcur = ctx->d_mpc_xcur; p.x0 = ctx->d_x0; p.babt = ctx->d_babt; p.rsq = ctx->d_rsq;
  p.A = ctx->d_plantA; p.Bm = ctx->d_plantB; p.b = ctx->d_plantb; p.plant_shared = plant_shared ? 1 : 0;
  p.sol_u = ctx->d_sol_u; p.iter = ctx->d_iter; p.status = ctx->d_status;
  p.x_traj = ctx->d_mpc_x; p.u_traj = ctx->d_mpc_u; p.iter_traj = ctx->d_mpc_iter; p.status_traj = ctx->d_mpc_status;
  const int grid = (ctx->B + 127) / 128;
  int rc = SRBD_OK;
  for (int t = 0; t < steps && rc == SRBD_OK; ++t) {
    p.t = t;
    mpc_embed_kernel<<<grid, 128, 0, ctx->stream>>>(p);
    ctx->launches++;
    // step 0 starts from the uploaded guess, later steps from the previous solution (the reference passes `solution`
    // back in, examples/example_mpc.cpp:114)
    ctx->warm_from_solution = ctx->args.warm_start && t > 0;
    rc = srbd_qp_solve(ctx);
    ctx->warm_from_solution = false;
    if (rc != SRBD_OK) break;
    mpc_plant_kernel<<<grid, 128, 0, ctx->stream>>>(p);
    ctx->launches++;
  }
  if (rc != SRBD_OK) return rc;
  CU(cudaGetLastError());
  if (x_traj) CU(cudaMemcpyAsync(x_traj, ctx->d_mpc_x, (size_t)(steps + 1) * B * nx * D, cudaMemcpyDeviceToHost, ctx->stream));
  if (u_traj) CU(cudaMemcpyAsync(u_traj, ctx->d_mpc_u, (size_t)steps * B * nu * D, cudaMemcpyDeviceToHost, ctx->stream));
  if (iter) CU(cudaMemcpyAsync(iter, ctx->d_mpc_iter, (size_t)steps * B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (status) CU(cudaMemcpyAsync(status, ctx->d_mpc_status, (size_t)steps * B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

// ---- SQP level ------------------------------------------------------------------------------------
int srbd_line_search(srbd_ctx* ctx) {
  if (int rc = require_srbd(ctx)) return rc;
  if (!ctx->solved) return fail(ctx, SRBD_ERR_STATE, "solve first");
  CU(cudaSetDevice(ctx->device));
  LsParams p{};
  p.B = ctx->B; p.N = ctx->L.N;
  p.mode = ctx->assembled_mode == SRBD_HARD_INEQ ? SRBD_HARD_INEQ : SRBD_BARRIER_SOFT;
  p.x = ctx->d_x; p.u = ctx->d_u; p.xref = ctx->d_xref; p.contact = ctx->have_contact ? ctx->d_contact : nullptr;
  p.dx = ctx->d_sol_x; p.du = ctx->d_sol_u; p.alpha = ctx->d_alpha; p.converged = ctx->d_conv; p.merit = ctx->d_merit;
  if (ctx->sqp_loop) {
    p.freeze = 1; p.run_gate = ctx->cur_gate; p.sqp_iter = ctx->d_sqp_iter;
    p.active_next = ctx->cur_gate ? const_cast<int*>(ctx->cur_gate) + 1 : ctx->d_active + 1;
  }
  line_search_kernel<<<(ctx->B + 3) / 4, 128, 0, ctx->stream>>>(p, ctx->d_model);
  ctx->launches++;
  CU(cudaGetLastError());
  ctx->traj_version++;   // the trajectory moved: the assembled QP no longer belongs to it
  return SRBD_OK;
}

int srbd_download_sqp_state(srbd_ctx* ctx, double* alpha, int* converged, double* merit) {
  if (int rc = require_srbd(ctx)) return rc;
  const size_t B = ctx->B;
  CU(cudaSetDevice(ctx->device));
  if (alpha) CU(cudaMemcpyAsync(alpha, ctx->d_alpha, B * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (converged) CU(cudaMemcpyAsync(converged, ctx->d_conv, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (merit) CU(cudaMemcpyAsync(merit, ctx->d_merit, B * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

int srbd_reset_sqp_state(srbd_ctx* ctx) {
  if (int rc = require_srbd(ctx)) return rc;
  std::vector<double> ones(ctx->B, 1.0);
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(ctx->d_alpha, ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_conv, 0, ctx->B * sizeof(int), ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

int srbd_sqp_iterate(srbd_ctx* ctx, int mode, int do_line_search) {
  if (int rc = srbd_linearize(ctx)) return rc;
  if (int rc = srbd_assemble(ctx, mode)) return rc;
  if (int rc = srbd_qp_solve(ctx)) return rc;
  if (do_line_search)
    if (int rc = srbd_line_search(ctx)) return rc;
  return SRBD_OK;
}

int srbd_sqp_solve(srbd_ctx* ctx, int mode, int max_iter) {
  if (int rc = require_srbd(ctx)) return rc;
  if (max_iter < 1 || max_iter > 1000) return fail(ctx, SRBD_ERR_ARG, "max_iter out of range");
  CU(cudaSetDevice(ctx->device));
  if (max_iter + 2 > ctx->active_alloc) {
    if (ctx->d_active) cudaFree(ctx->d_active);
    ctx->d_active = nullptr;
    CU(dalloc(&ctx->d_active, (size_t)max_iter + 2));
    ctx->active_alloc = max_iter + 2;
  }
  if (!ctx->d_sqp_iter) CU(dalloc(&ctx->d_sqp_iter, (size_t)ctx->B));
  CU(cudaMemsetAsync(ctx->d_active, 0, (size_t)(max_iter + 2) * sizeof(int), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_sqp_iter, 0, (size_t)ctx->B * sizeof(int), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_conv, 0, (size_t)ctx->B * sizeof(int), ctx->stream));
  // Iteration i runs gated on d_active[i] (the number of QPs iteration i-1 left unconverged; iteration 0 is not gated) and
  // its line search counts into d_active[i+1]: once every QP has converged the remaining launches return at once.  No
  // host round trip between the iterations; one read-back (srbd_download_sqp_state / srbd_download_sqp_iters) at the end.
  int rc = SRBD_OK;
  ctx->sqp_loop = true;
  for (int i = 0; i < max_iter && rc == SRBD_OK; ++i) {
    ctx->cur_gate = i == 0 ? nullptr : ctx->d_active + i;
    rc = srbd_sqp_iterate(ctx, mode, 1);
  }
  ctx->sqp_loop = false;
  ctx->cur_gate = nullptr;
  return rc;
}

int srbd_download_sqp_iters(srbd_ctx* ctx, int* iters) {
  if (int rc = require_srbd(ctx)) return rc;
  if (!iters) return SRBD_ERR_ARG;
  if (!ctx->d_sqp_iter) return fail(ctx, SRBD_ERR_STATE, "srbd_sqp_solve first");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(iters, ctx->d_sqp_iter, (size_t)ctx->B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

int srbd_solve_host_async(srbd_ctx* ctx, int mode, const double* x, const double* u, const double* xref,
                          const double* x0, const uint8_t* contact, double* sol_x, double* sol_u, int* iter,
                          int* status) {
  if (int rc = srbd_upload_traj(ctx, x, u, xref, x0, contact)) return rc;
  if (int rc = srbd_sqp_iterate(ctx, mode, 0)) return rc;
  const size_t B = ctx->B, S = ctx->L.N + 1, N = ctx->L.N, D = sizeof(double);
  if (sol_x) CU(cudaMemcpyAsync(sol_x, ctx->d_sol_x, B * S * 12 * D, cudaMemcpyDeviceToHost, ctx->stream));
  if (sol_u) CU(cudaMemcpyAsync(sol_u, ctx->d_sol_u, B * N * 12 * D, cudaMemcpyDeviceToHost, ctx->stream));
  if (iter) CU(cudaMemcpyAsync(iter, ctx->d_iter, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (status) CU(cudaMemcpyAsync(status, ctx->d_status, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  return SRBD_OK;
}

int srbd_wait(srbd_ctx* ctx) {
  if (!ctx) return SRBD_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

int srbd_solve_host(srbd_ctx* ctx, int mode, const double* x, const double* u, const double* xref, const double* x0,
                    const uint8_t* contact, double* sol_x, double* sol_u, int* iter, int* status) {
  if (int rc = srbd_upload_traj(ctx, x, u, xref, x0, contact)) return rc;
  if (int rc = srbd_sqp_iterate(ctx, mode, 0)) return rc;
  const size_t B = ctx->B, S = ctx->L.N + 1, N = ctx->L.N, D = sizeof(double);
  if (sol_x) CU(cudaMemcpyAsync(sol_x, ctx->d_sol_x, B * S * 12 * D, cudaMemcpyDeviceToHost, ctx->stream));
  if (sol_u) CU(cudaMemcpyAsync(sol_u, ctx->d_sol_u, B * N * 12 * D, cudaMemcpyDeviceToHost, ctx->stream));
  if (iter) CU(cudaMemcpyAsync(iter, ctx->d_iter, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (status) CU(cudaMemcpyAsync(status, ctx->d_status, B * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SRBD_OK;
}

int srbd_solve_host_graph(srbd_ctx* ctx, int mode, const double* x, const double* u, const double* xref, const double* x0,
                          const uint8_t* contact, double* sol_x, double* sol_u, int* iter, int* status) {
  if (int rc = require_srbd(ctx)) return rc;
  if (!x || !u || !xref || !x0) return fail(ctx, SRBD_ERR_ARG, "null trajectory pointer");
  CU(cudaSetDevice(ctx->device));
  const size_t B = ctx->B, S = ctx->L.N + 1, N = ctx->L.N, D = sizeof(double);
  const size_t nX = B * S * 12, nU = B * N * 12, n0 = B * 12;
  // staging layout (doubles first, then the two int arrays, then the contact bytes)
  const size_t oX = 0, oU = oX + nX, oR = oU + nU, o0 = oR + nX, oSX = o0 + n0, oSU = oSX + nX, oI = oSU + nU;
  const size_t bytes = oI * D + 2 * B * sizeof(int) + B * N * 2;
  if (!ctx->h_stage) CU(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_stage), bytes, cudaHostAllocDefault));
  double* hd = reinterpret_cast<double*>(ctx->h_stage);
  int* hi = reinterpret_cast<int*>(ctx->h_stage + oI * D);
  uint8_t* hc = reinterpret_cast<uint8_t*>(hi + 2 * B);
  std::memcpy(hd + oX, x, nX * D); std::memcpy(hd + oU, u, nU * D); std::memcpy(hd + oR, xref, nX * D);
  std::memcpy(hd + o0, x0, n0 * D);
  if (contact) std::memcpy(hc, contact, B * N * 2);
  const int has_contact = contact ? 1 : 0;
  if (!ctx->graph_exec || ctx->graph_mode != mode || ctx->graph_contact != has_contact) {
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    // one eager pass first: lazy allocations and function attributes must not happen inside a capture
    if (int rc = srbd_solve_host(ctx, mode, hd + oX, hd + oU, hd + oR, hd + o0, contact ? hc : nullptr, hd + oSX, hd + oSU, hi, hi + B))
      return rc;
    const long long l0 = ctx->launches;
    CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int rc = srbd_solve_host_async(ctx, mode, hd + oX, hd + oU, hd + oR, hd + o0, contact ? hc : nullptr, hd + oSX, hd + oSU, hi, hi + B);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc != SRBD_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(ctx, SRBD_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return fail(ctx, SRBD_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
    ctx->graph_launches = ctx->launches - l0;
    ctx->launches = l0;   // (the capture enqueued nothing)
    ctx->graph_mode = mode; ctx->graph_contact = has_contact;
  }
  CU(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
  ctx->launches += ctx->graph_launches;
  CU(cudaStreamSynchronize(ctx->stream));
  if (sol_x) std::memcpy(sol_x, hd + oSX, nX * D);
  if (sol_u) std::memcpy(sol_u, hd + oSU, nU * D);
  if (iter) std::memcpy(iter, hi, B * sizeof(int));
  if (status) std::memcpy(status, hi + B, B * sizeof(int));
  return SRBD_OK;
}

int srbd_fp64_peak(srbd_ctx* ctx, double* flops_per_s) {
  if (!ctx || !flops_per_s) return SRBD_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 4, threads = 256, iters = 4096;  // 32 warps per SM, ~3 ms per launch
  double* d_out = nullptr;
  CU(dalloc(&d_out, (size_t)blocks * threads));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0, ctx->stream));
    fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 1.0 + rep);
    ctx->launches++;
    CU(cudaEventRecord(e1, ctx->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 256.0 * 8.0 * (double)iters * blocks * (threads / 32) / (ms * 1e-3);  // 8 DMMA x 256 FMA per warp and trip
    if (rep > 0 && fl > best) best = fl;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *flops_per_s = best;
  return SRBD_OK;
}

}  // extern "C"
