// ipm_srbd.cuh — K3, throughput variant for QPs assembled by K2 (nx = nu = 12, 24 general rows on u only).
//
// Same algorithm, same constants and the same per-iteration sequence as ipm_solve.cuh (the generic kernel,
// which stays the path for arbitrary hpipm::OcpQp data); this variant exploits what K2 guarantees
// (NMPC_solver.cpp:300-309, SRBD_model.cpp:244-255):
//   * C = 0, S = 0, Q diagonal, and D = Ac is ONE constant 24x12 matrix made of two 12x6 blocks (one per
//     contact)                       -> Ac lives in shared memory; D^T Gamma D = sum_g Gamma_g W_g with the 21
//                                       lower-triangle products W_g of each row's 6-vector precomputed per CTA;
//   * only the lower side of the rows is active (ug masked), no box constraints, cold start.
// Mapping: independent warps, one QP per warp, persistent grid + atomic work counter; 6 warps per CTA and 2 CTAs per
// SM (SRBD_K3_WARPS x SRBD_K3_MIN_CTAS).  The register file allows 3 warps per SM sub-partition at 168 registers
// or 4 at 128 (nothing in between: 16384 registers per sub-partition), and the 4 x 128 build spills too much
// (445k vs 545k solves/s); among the 12-warp shapes the ones whose shared memory fits the 196 KB carve-out
// (6 x 2, 12 x 1) leave a 60 KB L1 instead of 28 KB and run 3-4 % faster than 4 x 3.
// The solve runs on the FP64 tensor cores (mma.sync m8n8k4 f64, "DMMA"): one
// instruction is 256 FMAs, i.e. 8 warp-wide DFMAs plus their shared-memory operand fetches in ONE issue slot
// (scripts/microbench/fp64_pipes.cu: DMMA and DFMA share the FP64 pipe on B200, 36.9 vs 35.8 TFLOP/s, so the
// gain is issue slots, not peak).  The stage matrix lives in the registers as 8x4 FRAGMENTS
//     F[I][p] : lane (r = lane>>2, t = lane&3) holds M[8I + r][4p + t]
// which is at the same time the A-operand layout (row r, k = t), the B-operand layout of the transposed tile and,
// when the B operand's rows are permuted by pi(r) = (r>>1) + 4(r&1), the two halves of the accumulator tile:
//     D(8x8) = A(8x4) . B(4x8)^T-rows-permuted  ->  c0 = F[I][2J], c1 = F[I][2J+1].
// So AL = G P, M = H + AL G^T and the trailing updates of the blocked (4-column BLASFEO-style panels) Cholesky
// chain without any register shuffles: 18 + 18 + 12 DMMA per stage.  The 4-column panels themselves are solved
// row-per-lane through a [37][4] shared-memory panel (every lane factors the 4x4 diagonal block redundantly, then
// substitutes its own row); four appended identity rows per panel turn the same substitution into L_pp^-T, the
// diagonal blocks of the blocked triangular solves of the vector sweeps (substitution, not an explicit 12x12
// inverse: the residual of the Newton step decides convergence at tol 1e-8).
// The gradient recursion (the "+1 row" of potrf_l_mn) and the vector sweeps (S2/S5 forward rollout fused with
// dlam/dt, the step length and the mu_aff sums; S4 vector-only backward with the centering correction on the fly;
// S6 variable update fused with the residuals) run in VECTOR-FRAGMENT form: a 12-vector is three registers valid
// in lanes 0..3 = row 0 of an A operand, y = A x is 2 x 3 DMMA with A as pi-permuted B fragments, and the
// accumulator pair of output tile I is k-tiles 2I, 2I+1 of y: chained products stay in registers; only the 24
// constraint rows are handled row-per-lane (one shared-memory round trip per stage).
// Memory: every sweep is SOFTWARE PIPELINED: while stage k is computed, the tiles of the next stage (BAbt record,
// P / factor panels, the R block of RSQrq) stream into the other half of a shared-memory double buffer with
// cp.async, and the next stage's per-row vectors are prefetched into registers (raw loads only: arithmetic on a
// prefetched value would wait for it at the prefetch point), so HBM/L2 latency overlaps the FP64 work.
// Second half of round 2: the residual sweep is fused backward into the factorization sweep (sweep_resfac: four sweeps per
// iteration instead of five), and the throughput instantiation streams COMPACT BAbt records (kCG): the 288 model constants
// of a record stay resident in the shared-memory tiles (filled once per warp by two TMA bulk copies), each stage copies the
// 36 stage-dependent 16-byte chunks of K1's dyn record over them; the row masks follow from the assemble mode.  All of it
// bit-identical to the dense path (SRBD_K3_CG=0); DRAM traffic per QP 8.6 -> 5.3 MB (DESIGN.md sections 3 and 5.0).
#pragma once
#include <cuda_runtime.h>

#include "ipm_solve.cuh"
#include "srbd_model.cuh"

namespace srbd {

struct SrbdIpmParams {
  int B, N;
  srbd_ipm_args a;
  const double* babt;   // [B][N][336]
  int asm_mode;         // kCG instantiations: the mode srbd_assemble ran in (the row masks follow from it)
  const double* gconst; // [336] dense BAbt record of (QP 0, stage 1) (K1): the model constants, kCG instantiations
  const double* gdyn;   // [B][N][kBabtDyn] stage-dependent chunks of the same records (K1; layout.cuh), kCG instantiations
  const double* srec;   // [B][N+1][kSrec] compact stage records of K2 (srbd_model.cuh: R tile, gradient row, lg, masks)
  const double* x0;     // [B][12]
  const ModelDev* model;
  double* ws;           // [gridDim.x*4][ws_size]
  int ws_size;
  int* counter;
  double *sol_x, *sol_u, *sol_pi, *sol_lam, *sol_t;
  int* iter;
  int* status;
  double* res_max;
  srbd_batch_stats* bstats;
  // rescue list (or null): QPs that run to iter_max are appended here instead of being counted; capi.cu then runs
  // the generic kernel (row-by-row substitution, no block inverses) on exactly those
  int* retry_list;
  int* retry_count;
  // optional work list (the first rescue stage, capi.cu): solve only the QPs qlist[0 .. *qcount)
  const int* qlist;
  const int* qcount;
  // device-side dispatch (QP-level uploads, capi.cu): run only if *gate == gate_value
  const int* gate;
  int gate_value;
  // device-side SQP loop (srbd_sqp_solve): return at once if *run_gate == 0; skip QPs whose frozen[] flag is set (their NMPC
  // problem has converged: the reference leaves its SQP loop then, NMPC_solver.cpp:372-374)
  const int* run_gate;
  const int* frozen;
  // optional exports (QP-level uploads through the hpipm-cpp facade, which fills them like the reference does,
  // hpipm-cpp/src/ocp_qp_ipm_solver.cpp:337-387): the Riccati matrices of the last factorization, pi[0], and the
  // per-iteration statistics table [B][stat_rows][18]
  double *ric_P, *ric_p, *ric_K, *ric_k;   // null: none
  double* ric_Lr0;      // optional (with ric_P): [B][144] column-major Cholesky factor Lr of stage 0 (d_ocp_qp_ipm_get_ric_Lr(.., 0, ..))
  const double* raw0;   // [B][raw0_stride] raw stage-0 blocks A0 | B0 | b0 | S0 | Q0 | q0 (column-major), for the stage-0 export
  int raw0_stride;
  double* stat;         // null: none
  int stat_rows;
};

namespace v2 {
#ifndef SRBD_K3_WARPS
#define SRBD_K3_WARPS 6       // warps (= QPs in flight) per CTA
#endif
#ifndef SRBD_K3_MIN_CTAS
#define SRBD_K3_MIN_CTAS 2    // resident CTAs per SM the register allocation is bounded for
#endif
// ---- compile-time knobs of the A/B experiments (defaults = what measured fastest) ------------
// stage loops unrolled by two: the loop-carried copies of the register-prefetched vectors (cur = nxt, 22-32 moves
// per stage) and the buffer selects disappear, but every unrolled sweep costs 5 % (instruction cache): off
#ifndef SRBD_K3_UNROLL_FAC
#define SRBD_K3_UNROLL_FAC 0
#endif
#ifndef SRBD_K3_UNROLL_BWD
#define SRBD_K3_UNROLL_BWD 0
#endif
#ifndef SRBD_K3_UNROLL_FWD
#define SRBD_K3_UNROLL_FWD 0
#endif
#ifndef SRBD_K3_UNROLL_RES
#define SRBD_K3_UNROLL_RES 0
#endif
// two half-used output tiles that multiply the same vector share one DMMA (bit-identical results)
#ifndef SRBD_K3_MERGE_HALVES
#define SRBD_K3_MERGE_HALVES 1
#endif
#ifndef SRBD_K3_DUMMY_DMMA
#define SRBD_K3_DUMMY_DMMA 0   // experiment only
#endif
#ifndef SRBD_K3_L2PF
#define SRBD_K3_L2PF 0    // off (measured slower); bit 0: vectors, 1: P / factor panels, 2: BAbt record, 3: stage record
#endif
// QBASE: per-QP base pointers of the packed QP data held opaque (+1 %).  WBASE: the per-lane workspace bases as well
// (fewer address instructions, but 116 B of spills: -3 %)
#ifndef SRBD_K3_QBASE
#define SRBD_K3_QBASE 1
#endif
#ifndef SRBD_K3_WBASE
#define SRBD_K3_WBASE 0
#endif
// Tile prefetch engine PER SWEEP (bit mask: 1 factorization S1, 2 backward vector sweep S4, 4 forward sweep S2,
// 8 residual sweep S6).  Bit set: TMA bulk copies (cp.async.bulk.shared.global, completion on a per-warp mbarrier pair)
// issued by ONE lane, one instruction per contiguous block (the BAbt record, the [P | factor panels] tile, the R tile),
// nothing through the LSU.  Bit clear: Ampere-style per-lane 16-byte cp.async (LDGSTS), the round-1 engine.
// Measured (16384 QPs burst / 65536 QPs sustained, profiles/r2_k3_tile_engine_ab.txt): every sweep on bulk copies (nine
// per stage with the padded BAbt tile) 483 k / 496 k solves/s against 588 k / 546 k for cp.async; three per stage (unpadded
// tile) 549 k / 548 k; S1 + S6 only 554 k / 530 k; S1 only 568 k / 534 k; S6 only 575 k / 540 k against 580 k / 545 k
// without any.  A bulk copy has a longer latency than the per-lane copies and the prefetch distance is one stage (the
// shared memory is full), which the short vector sweeps S2 / S4 (2.6-3.0 k cycles per stage) and the factorization cannot
// hide; the residual sweep has no recursion to wait for and takes them at no measurable cost when the GPU is full.  A
// single QP alone (BASELINE config 5) has nothing to overlap the longer latency with: 3.75 ms against 3.59 ms per N = 50
// solve.  Hence two instantiations: SRBD_K3_TMA (default: S6) for batches that fill the machine, SRBD_K3_TMA_SMALL
// (default: none) for batches of fewer QPs than SMs (capi.cu picks).
// Since the residual sweep is fused into the factorization sweep (SRBD_K3_FUSE) the stand-alone S6 no longer runs in the
// throughput instantiations, so with the default mask (8) their per-stage streams are all cp.async; re-measured after the
// fusion: bit 1 (the R tile of the fused sweep as one bulk copy) 599.9 k against 608.9 k sustained
// (profiles/r2c_k3_variants_ab.txt).  A non-zero mask still sets up the per-warp mbarriers, which the one-time constant
// fill of the compact BAbt tiles uses (fill_G_constants: two whole-record bulk copies per warp and kernel).
#ifndef SRBD_K3_TMA
#define SRBD_K3_TMA 8
#endif
// residual sweep fused into the factorization sweep (sweep_resfac), throughput instantiations only
#ifndef SRBD_K3_FUSE
#define SRBD_K3_FUSE 1
#endif
#ifndef SRBD_K3_TMA_SMALL
#define SRBD_K3_TMA_SMALL 0
#endif
// panel stride of the BAbt tile in shared memory.  48 = the record as it lies in HBM (one bulk copy); 54 (= 2 mod 4) makes
// the row-permuted B fragments bank-conflict free but measures no faster (580 k vs 576 k burst, 545 k vs 538 k sustained)
#ifndef SRBD_K3_GP
#define SRBD_K3_GP 48
#endif
// experiments on the inverse Cholesky pivots of EVERY instantiation (1: 1 / sqrt, 2: CUDA rsqrt(), 3: one more Newton step);
// the product uses the template parameter kPivot of SrbdSolver instead
#ifndef SRBD_K3_EXACT_RSQRT
#define SRBD_K3_EXACT_RSQRT 0
#endif
constexpr int kWarps = SRBD_K3_WARPS;
constexpr int kMinCtas = SRBD_K3_MIN_CTAS;
// per-stage workspace block (doubles)
constexpr int oZ = 0, oDZ = 24, oRG = 48, oLAM = 72, oT = 96, oDLAM = 120, oDT = 144, oRD = 168, oRM = 192, oRMB = 216,
              oPI = 240, oDPI = 252, oRB = 264, oPV = 276, oLV = 288, oP = 300, oFT = 444, oPRB = 750, kStage = 762;
// factor tile of a stage: 3 column panels x [25 rows][4]: rows 0..11 = rows of L^-T (E rows), 12..23 = Ls, 24 = lv
// team (latency) mode only: a second (z, pi) slot per stage BEHIND the N + 1 stage blocks (the stage blocks of the throughput
// mode are untouched).  The team's residual sweep reads iterate slot s and writes the updated one to slot s ^ 1: a warp may
// then read its neighbours' z_{k+1}, pi_{k-1} while they store theirs.
constexpr int kAlt = 36;           // [z 24 | pi 12]
constexpr int kPanF = 102;         // panel stride in the factor tile: rows 0..24 (100) + 2 (bank-conflict-free fragments)
constexpr int kPan = 148;          // the full panel during the factorization: + rows 25..36 = rows 0..11 of L
// shared memory (doubles)
constexpr int kGP = SRBD_K3_GP;    // panel stride of the BAbt tile (4 rows x 12 cols [+ 6 padding])
constexpr int kW2 = 22;            // row stride of W (21 lower-triangle products of a constraint row's 6-vector)
constexpr int sAC = 0;             // [24][12] constraint Jacobian, row-major
constexpr int sW = sAC + 288;      // [24][22]
constexpr int sQ = sW + 24 * kW2;  // diag(Q) (12), R (1), pad
constexpr int kCtaShared = sQ + 32;    // 848 doubles = 53 x 128 B (sQ + 16 .. : one work-counter int per warp)
// every tile starts on a 128-byte line (16 doubles): a 512-byte cp.async instruction then writes 4 wavefronts, not 5
constexpr int kGT = 384;           // one BAbt tile: 7 x 54 = 378 -> 384
constexpr int kFT = 464;           // [P 144 | factor panels 3 x 102 = 306] = 450 -> 464
constexpr int kRT = 144;           // [R lower-panel prefixes 96 | Q diag 12 | rq row 24] = 132 -> 144
constexpr int wG0 = 0, wG1 = kGT;
constexpr int wF0 = 2 * kGT, wF1 = wF0 + kFT;
// the factorization / residual sweeps do not use the factor tiles: the running P_{k+1}, the three Cholesky panels
// and both R tiles live in that region
constexpr int wP = wF0;            // [12][12]                   144 -> 160
constexpr int wPAN = wF0 + 160;    // 3 x [37][4]                444 -> 608
constexpr int wR0 = wF0 + 608, wR1 = wR0 + kRT;
static_assert(wR1 + kRT <= wF0 + 2 * kFT, "R tiles do not fit behind the factorization scratch");
constexpr int wS = wF0 + 2 * kFT;  // 44 -> 48
constexpr int wQX = wS + 48;       // 24 Gamma
constexpr int wqx = wQX + 24;      // 24 gamma
constexpr int wSG = wqx + 24;      // 24 gradient
constexpr int wSX = wSG + 24;      // 24 z of the current stage
constexpr int wXN = wSX + 24;      // 12 x_{k+1} / pi_k
constexpr int wT = wXN + 12;       // 12 t / lv
constexpr int wPV = wT + 12;       // 12 p_{k+1}
constexpr int wDI = wPV + 12;      // 12 (spare)
constexpr int wLAM = wDI + 12;     // 24 lam (residual sweep)
constexpr int wBAR = wLAM + 24;    // 2 mbarriers (TMA tile engine) in the padding
constexpr int kWarpShared = wLAM + 24 + 8;   // multiple of 16
static_assert(kWarpShared % 16 == 0 && kCtaShared % 16 == 0, "tiles must start on 128-byte lines");
constexpr int kSmemBytes = (kCtaShared + kWarps * kWarpShared) * 8;
}  // namespace v2

// load of a workspace double (private to the warp, re-read once per sweep: no reuse in L1).  SRBD_K3_WSLD selects the
// cache operator for A/B runs: 0 = ld.global.cg (L2 only), 1 = L1::no_allocate, 2 = default (.ca), 3 = .cs (streaming):
// 574 / 567 / 578 / 574 k solves/s, i.e. no measurable effect
#ifndef SRBD_K3_WSLD
#define SRBD_K3_WSLD 0
#endif
__device__ __forceinline__ double ws_ld(const double* p) {
#if SRBD_K3_WSLD == 0
  return __ldcg(p);
#elif SRBD_K3_WSLD == 1
  double v;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];\n" : "=d"(v) : "l"(p));
  return v;
#elif SRBD_K3_WSLD == 2
  return *reinterpret_cast<const volatile double*>(p);
#else
  return __ldcs(p);
#endif
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
// L2 prefetch of one 128-byte line per lane (no destination register), two stages ahead of the sweeps, meant to let the
// register prefetches and cp.async copies of the next stage find their lines in L2 (11 % of the stall samples of v13 are
// the loop-top wait for them).  EXPERIMENT, off by default: with 1776 QPs in flight the working set is twice the L2,
// and the deeper prefetch evicts more than it saves (519-551 k against 567 k solves/s)
__device__ __forceinline__ void l2_prefetch(const void* p) {
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// ---- TMA bulk copies + mbarrier (sm_90+; SASS: UBLKCP / SYNCS) -------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SRBD_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SRBD_MBAR_DONE;\n"
      "bra SRBD_MBAR_WAIT;\n"
      "SRBD_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16), completes `bytes` transactions on `bar`
__device__ __forceinline__ void bulk_g2s(double* smem_dst, const double* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// generic-proxy accesses of this thread (and, after a __syncwarp, of its warp) before async-proxy writes of the same memory
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// D(8x8) = A(8x4) * B(4x8) + C on the FP64 tensor cores; fragment layout in the header comment
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
      : "=d"(d0), "=d"(d1)
      : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
// kTma: the tile-engine mask of this instantiation (SRBD_K3_TMA above).  kPivot: how the inverse Cholesky pivots are formed:
// 0 = MUFU seed + one third-order step (1.25 ulp, the throughput flavour), 1 = CUDA's rsqrt() (1 ulp, 5 % slower).  Both
// are faithful; they ROUND differently.  About 2-4 QPs per million sit on a knife edge of the IPM's rounding floor and run
// to iter_max in one flavour and not in the other (scripts/count_iter_max.py: 2 + 1 per 2 x 1048576 QPs with flavour 0,
// 1 + 1 OTHER ones with flavour 1, 4 + 0 with an extra Newton step): the first rescue stage re-solves the listed QPs
// with flavour 1 at tensor-core speed (1.4 ms for one QP instead of 14-17 ms in the generic kernel).
// kTeam: 0 = one warp per QP (throughput); n > 0 = LATENCY mode, the n warps of the CTA work on ONE QP: the residual sweep
// (no recursion: a fifth of a solve) is split over them by stage, the four Riccati recursions stay on warp 0.  Results are
// bit-identical to kTeam = 0 (same operations per element; the duality measure is summed in the original order).
// kExp: the instantiation that can write the facade's exports (Riccati matrices, pi[0], statistics table); the throughput
// instantiations carry none of that code.
// kCG: COMPACT BAbt streaming (QPs linearized by K1 only).  288 of the 336 doubles of a BAbt record are model constants (0, 1,
// dt, dt/m), the same for every stage of every QP: the warp fills both of its BAbt tiles with them ONCE (from any dense
// record of a stage >= 1) and then, per stage and sweep, copies only the 36 16-byte chunks that hold a stage-dependent
// element from K1's dyn record (layout.cuh: babt_dyn_off) over them: 576 instead of 2304 / 2688 bytes per stage and sweep,
// one or two cp.async per lane instead of six or seven.  The tile is never written by anything else.  Stage 0 takes the
// same path: its dyn record carries b0 at row 24 (the dense stage-0 record has it at row 12) and finite A^T entries in
// rows 12..23, which every sweep either masks (xr) or multiplies by dx0 = 0 -- the same exact zeros as before.
template <int kTma, int kPivot, int kTeam, bool kExp, bool kCG = false>
struct SrbdSolver {
  const SrbdIpmParams& p;
  int lane, q, N;
  double* W;          // workspace of this warp
  const double* cAc;  // shared: Ac [24][12]
  const double* cW;   // shared: W [24][22]
  const double* cQ;   // shared: diag(Q) (12), R
  double* sm;         // this warp's shared block
  const double* sG;   // current BAbt tile
  const double* sF;   // current factor tile
  const double* sR;   // current R tile
  int wid;            // warp of the CTA (team mode: 0 = leader)
  int zs;             // team mode: the slot holding the current (z, pi)
  double* cred;       // CTA-shared scratch of the team: [kTeam][6] partial norms | broadcast slots
  // fragment coordinates of this lane (see the header comment)
  int fr, ft, fpi;
  int offS[5];        // index into the 42 D^T Gamma D sums of this lane's element of the u-block fragments, or -1
  int gdo;            // kCG: tile offset of dyn chunk `lane`

  __device__ SrbdSolver(const SrbdIpmParams& p_, double* cta, double* warp_sm, int warp_global)
      : p(p_), lane(threadIdx.x & 31), q(0), N(p_.N) {
    W = p.ws + (size_t)warp_global * p.ws_size;
    cAc = cta + v2::sAC;
    cW = cta + v2::sW;
    cQ = cta + v2::sQ;
    sm = warp_sm;
    wid = threadIdx.x >> 5;
    zs = 0;
    cred = cta + v2::kCtaShared;
    sG = sm; sF = sm + v2::wF0; sR = sm + v2::wR0;
    fr = lane >> 2; ft = lane & 3; fpi = (fr >> 1) + 4 * (fr & 1);
    gdo = tile_off(babt_dyn_off(lane));
#if SRBD_K3_WBASE
    Wc = W + (lane < 24 ? lane : 0);
    Wf = W + ft;
    asm volatile("" : "+l"(Wc), "+l"(Wf));
#endif
    // u-block fragments in the order (I,p) = (0,0) (0,1) (1,0) (1,1) (1,2)
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      const int I = f < 2 ? 0 : 1, pp = f < 2 ? f : f - 2;
      const int i = 8 * I + fr, c = 4 * pp + ft;
      int o = -1;
      if (i < 12 && c <= i && (i / 6) == (c / 6)) {
        const int leg = i / 6, ii = i - 6 * leg, cc = c - 6 * leg;
        o = 21 * leg + ii * (ii + 1) / 2 + cc;
      }
      offS[f] = o;
    }
  }
  __device__ __forceinline__ double* ws(int k, int off) const { return W + (size_t)k * v2::kStage + off; }
#if SRBD_K3_WBASE
  double *Wc, *Wf;    // W + min(lane, 23)-style index lc / W + (lane & 3): opaque per-lane bases of the vector loads
  __device__ __forceinline__ double* wsc(int k, int off) const { return Wc + k * v2::kStage + off; }
  __device__ __forceinline__ double* wsf(int k, int off) const { return Wf + k * v2::kStage + off; }
#else
  __device__ __forceinline__ double* wsc(int k, int off) const { return ws(k, off) + (lane < 24 ? lane : 0); }
  __device__ __forceinline__ double* wsf(int k, int off) const { return ws(k, off) + ft; }
#endif
  // (z, pi) of stage k in slot s (s is 0 outside the team mode): per-lane pointers like wsc / wsf
  __device__ __forceinline__ double* zc(int k, int s) const {
    return (kTeam && s) ? wsc(N + 1, 0) + k * v2::kAlt : wsc(k, v2::oZ);
  }
  __device__ __forceinline__ double* pic(int k, int s) const {
    return (kTeam && s) ? wsc(N + 1, 0) + k * v2::kAlt + 24 : wsc(k, v2::oPI);
  }
  __device__ __forceinline__ double* pif(int k, int s) const {
    return (kTeam && s) ? wsf(N + 1, 0) + k * v2::kAlt + 24 : wsf(k, v2::oPI);
  }
  // per-QP, per-lane base pointers of the packed QP data (set once per solve; kept opaque so that the compiler
  // holds / reloads them instead of re-deriving them from (q, N, lane) with 64-bit multiplies at every stage)
  static constexpr int kLaneOff = 2;   // every lane copies its own 16-byte chunk (lane 0's pointer is the record base)
#if SRBD_K3_QBASE
  const double *qG, *qT;
  __device__ __forceinline__ void set_qp(int qp) {
    q = qp;
    qG = kCG ? p.gdyn + (size_t)q * N * kBabtDyn + kLaneOff * lane : p.babt + (size_t)q * N * 336 + kLaneOff * lane;
    qT = p.srec + (size_t)q * (N + 1) * kSrec + kLaneOff * lane;
    asm volatile("" : "+l"(qG), "+l"(qT));
  }
  __device__ __forceinline__ const double* gDynL(int k) const { return qG + k * kBabtDyn; }   // + 2 * lane
  __device__ __forceinline__ const double* gBAbtL(int k) const { return qG + k * 336; }   // + 2 * lane
  __device__ __forceinline__ const double* gRecL(int k) const { return qT + k * kSrec; }  // + 2 * lane
#else
  __device__ __forceinline__ void set_qp(int qp) { q = qp; }
  __device__ __forceinline__ const double* gBAbtL(int k) const { return p.babt + ((size_t)q * N + k) * 336 + kLaneOff * lane; }
  __device__ __forceinline__ const double* gDynL(int k) const { return p.gdyn + ((size_t)q * N + k) * kBabtDyn + kLaneOff * lane; }
  __device__ __forceinline__ const double* gRecL(int k) const {
    return p.srec + ((size_t)q * (N + 1) + k) * kSrec + kLaneOff * lane;
  }
#endif
  // lg / mask of row min(lane, 23)-ish (lanes >= 24 read row 0) from the stage record
  __device__ __forceinline__ const double* gDL(int k) const {
    return gRecL(k) + 144 + (lane < 24 ? lane : 0) - kLaneOff * lane;
  }
  __device__ __forceinline__ const double* gMaskL(int k) const { return gDL(k) + 24; }
  // mask of this lane's constraint row.  QPs assembled by K2 (the kCG instantiation): a function of the assemble mode and the
  // row alone (srbd_model.cuh: 20 hard rows per stage in HARD_INEQ mode -- all but the x^T tau pair of each leg --, none in
  // BARRIER_SOFT mode), so nothing is loaded; otherwise from the stage record
  __device__ __forceinline__ double mask_of(int k) const {
    if (kCG) {
      const int g = lane < 24 ? lane : 0, rr = g >= 12 ? g - 12 : g;
      return (p.asm_mode == SRBD_HARD_INEQ && rr < 10) ? 1.0 : 0.0;
    }
    return __ldg(gMaskL(k));
  }

  // ---- asynchronous tile prefetch (cp.async, no registers) -------------------------------------------
  // BAbt record (panels of 4 rows x 12 = 48 doubles) -> padded panels; np = 6: rows 0..23, 7: + the b row.
  // One 24-lane instruction per panel: every address is base + immediate.
  // per-warp mbarrier pair (one per tile buffer) in the padding of the warp's shared block; tph: bit b = the phase
  // parity the next wait on barrier b expects
  unsigned tph;
  __device__ __forceinline__ unsigned long long* bar(int b) const {
    return reinterpret_cast<unsigned long long*>(sm + v2::wBAR) + b;
  }
  __device__ __forceinline__ void tiles_init() {
    tph = 0;
    if (kTma) {
      if (lane == 0) {
        mbar_init(bar(0), 1);
        mbar_init(bar(1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        fence_proxy_async();
      }
      __syncwarp();
    }
  }
  // Stage k's tiles have landed in buffer b.  TMA: lane 0 issued every expect_tx / copy of this phase before (program
  // order), its arrival closes the phase's arrival count; the phase completes when the bytes are in.
  template <bool TMA>
  __device__ __forceinline__ void tiles_wait(int b) {
    if (TMA) {
      if (kCG) cp_async_wait_all();   // (the dyn chunks of the BAbt tile travel by cp.async in every sweep)
      if (lane == 0) mbar_arrive(bar(b));
      mbar_wait(bar(b), (tph >> b) & 1u);
      tph ^= 1u << b;
    } else {
      cp_async_wait_all();
    }
  }
  // TMA: before the first copy into a buffer the warp has read / written with ordinary shared-memory instructions (call
  // after the __syncwarp that ends those accesses)
  template <bool TMA>
  __device__ __forceinline__ void tiles_begin() const {
    if (TMA && lane == 0) fence_proxy_async();
  }
  // BAbt record (panels of 4 rows x 12 = 48 doubles) -> panels of stride kGP; np = 6: rows 0..23, 7: + the b row.
  template <bool TMA>
  __device__ __forceinline__ void prefetch_G(int k, int b, int np = 6) {
    double* dst = sm + (b ? v2::wG1 : v2::wG0);
    if (kCG) {   // the stage-dependent chunks only (chunks 0..11 = the b row: np == 7)
      const double* dyn = gDynL(k);   // + 2 * lane = chunk `lane`
      if (np == 7 || lane >= 12) cp_async16(dst + gdo, dyn);
      // chunks 32..35: record offsets 210, 214 (panel 4) and 252, 256 (panel 5)
      if (lane < 4) cp_async16(dst + (lane < 2 ? 210 + 4 * (v2::kGP - 48) : 244 + 5 * (v2::kGP - 48)) + 4 * lane, dyn + 64);
      return;
    }
    const double* src = gBAbtL(k);   // + 2 * lane
    if (TMA) {
      if (lane == 0) {
        mbar_expect_tx(bar(b), np * 384);
        if (v2::kGP == 48) {
          bulk_g2s(dst, src, np * 384, bar(b));
        } else {
#pragma unroll
          for (int pnl = 0; pnl < 7; ++pnl)
            if (pnl < np) bulk_g2s(dst + pnl * v2::kGP, src + pnl * 48, 384, bar(b));
        }
      }
    } else if (lane < 24) {   // one 24-lane instruction per panel: every address is base + immediate
#pragma unroll
      for (int pnl = 0; pnl < 7; ++pnl)
        if (pnl < np) cp_async16(dst + 2 * lane + pnl * v2::kGP, src + pnl * 48);
    }
  }
  // [P_{k+1} (144) |] factor panels of stage k (3 x 102): contiguous in the workspace (P_{k+1} lives in block k)
  template <bool TMA>
  __device__ __forceinline__ void prefetch_F(int k, int b, bool with_P) {
    double* dst = sm + (b ? v2::wF1 : v2::wF0);
    const double* src = ws(k, v2::oP);
    if (TMA) {
      if (lane == 0) {
        if (with_P) {
          mbar_expect_tx(bar(b), 450 * 8);
          bulk_g2s(dst, src, 450 * 8, bar(b));
        } else {
          mbar_expect_tx(bar(b), 306 * 8);
          bulk_g2s(dst + 144, src + 144, 306 * 8, bar(b));
        }
      }
    } else {
      dst += 2 * lane; src += 2 * lane;
      if (with_P) {
#pragma unroll
        for (int i = 0; i < 2; ++i) cp_async16(dst + 64 * i, src + 64 * i);
        if (lane < 8) cp_async16(dst + 128, src + 128);
      }
      // chunks 72 .. 224
#pragma unroll
      for (int i = 0; i < 4; ++i) cp_async16(dst + 144 + 64 * i, src + 144 + 64 * i);
      if (lane < 25) cp_async16(dst + 144 + 256, src + 144 + 256);
    }
  }
  // R tile of stage k (the lower 12 x 12 block of rows 0..11 of RSQrq as panel prefixes, 96 doubles): the head of the
  // compact stage record, a linear copy.  R_k is NOT a constant: the rows that stay a relaxed barrier in HARD_INEQ
  // mode add Ac^T diag(b'') Ac (NMPC_solver.cpp:308)
  template <bool TMA>
  __device__ __forceinline__ void prefetch_Rblk(int k, int b) {
    const double* src = gRecL(k);    // + 2 * lane
    double* dst = sm + (b ? v2::wR1 : v2::wR0);
    if (TMA) {
      if (lane == 0) {
        mbar_expect_tx(bar(b), 96 * 8);
        bulk_g2s(dst, src, 96 * 8, bar(b));
      }
    } else {
      dst += 2 * lane;
      cp_async16(dst, src);
      if (lane < 16) cp_async16(dst + 64, src + 64);
    }
  }
  // ... and with the gradient row n = [r; q] (record offsets 108..131)
  template <bool TMA>
  __device__ __forceinline__ void prefetch_R(int k, int b) {
    const double* src = gRecL(k);
    double* dst = sm + (b ? v2::wR1 : v2::wR0);
    if (TMA) {
      if (lane == 0) {
        mbar_expect_tx(bar(b), 132 * 8);
        bulk_g2s(dst, src, 132 * 8, bar(b));
      }
    } else {
      dst += 2 * lane;
      cp_async16(dst, src);
      cp_async16(dst + 64, src + 64);
      if (lane < 2) cp_async16(dst + 128, src + 128);
    }
  }
  // Lines of stage k that a sweep will read: the vector part of the workspace block (doubles 0..299), optionally
  // P_{k+1} (300..443) and the factor panels / P rb (444..761), the BAbt record and the compact stage record.
  // Every lane takes one line per instruction (the blocks are not line aligned: one spare line each).
  __device__ __forceinline__ void prefetch_L2(int k, bool with_P, bool with_F, bool with_rec) const {
#if SRBD_K3_L2PF
    const double* w = ws(k, 0) + 16 * lane;
    if ((SRBD_K3_L2PF & 1) && lane < 20) l2_prefetch(w);                       // vectors: 2400 B = 19 (+1) lines
    if ((SRBD_K3_L2PF & 2) && with_P && lane < 10) l2_prefetch(w + 304);       // P: 1152 B
    if ((SRBD_K3_L2PF & 2) && with_F && lane < 21) l2_prefetch(w + 448);       // panels + P rb: 2544 B = 20 (+1) lines
    if ((SRBD_K3_L2PF & 4) && lane < 22) l2_prefetch(gBAbtL(k) + 14 * lane);   // BAbt: 2688 B = 21 (+1) lines
    if ((SRBD_K3_L2PF & 8) && with_rec && lane < 12) l2_prefetch(gRecL(k) + 14 * lane);  // stage record: 12 lines
#endif
  }
  // kCG: the constants of a BAbt record into both tiles, once per warp, from a dense record of any stage >= 1 (its
  // stage-dependent elements are overwritten by the dyn chunks of every stage before the tile is read)
  // (one whole contiguous record per tile, once per warp and kernel: the natural job for a TMA bulk copy -- one instruction
  // per tile from one lane, completion on the warp's mbarrier; the per-stage streams stay on per-lane cp.async, which
  // measured faster in every sweep, see SRBD_K3_TMA above)
  // offset in the BAbt tile (panels of stride kGP) of offset o of the dense record (panels of 48)
  static __device__ __forceinline__ int tile_off(int o) { return v2::kGP == 48 ? o : (o / 48) * v2::kGP + o % 48; }
  __device__ __forceinline__ void fill_G_constants(const double* dense_rec) {
    if (kTma) {
      if (lane == 0) {
        mbar_expect_tx(bar(0), 2 * 336 * 8);
        if (v2::kGP == 48) {
          bulk_g2s(sm + v2::wG0, dense_rec, 336 * 8, bar(0));
          bulk_g2s(sm + v2::wG1, dense_rec, 336 * 8, bar(0));
        } else {
          for (int pnl = 0; pnl < 7; ++pnl) {
            bulk_g2s(sm + v2::wG0 + pnl * v2::kGP, dense_rec + pnl * 48, 384, bar(0));
            bulk_g2s(sm + v2::wG1 + pnl * v2::kGP, dense_rec + pnl * 48, 384, bar(0));
          }
        }
      }
      tiles_wait<true>(0);
    } else {
      for (int e = lane; e < 168; e += 32) {
        cp_async16(sm + v2::wG0 + tile_off(2 * e), dense_rec + 2 * e);
        cp_async16(sm + v2::wG1 + tile_off(2 * e), dense_rec + 2 * e);
      }
      cp_async_wait_all();
    }
    __syncwarp();
  }
  __device__ __forceinline__ void set_bufs(int b) {
    sG = sm + (b ? v2::wG1 : v2::wG0);
    sF = sm + (b ? v2::wF1 : v2::wF0);
    sR = sm + (b ? v2::wR1 : v2::wR0);
  }

  // ------------------------------------------------------------------------------------------------
  // S1: backward Riccati factorization sweep
  // ------------------------------------------------------------------------------------------------
  // The whole stage runs in fragment form: the gradient recursion (the "+1 row" of potrf_l_mn) is the S4 body on
  // the same fragments (P rb, D^T gamma, G t as DMMA gemvs; its trailing updates reuse the B fragments of the
  // matrix update); Q = diag(Q) is a model constant for k < N (NMPC_solver.cpp:305), R_k is read from RSQrq.
  struct S1v { double mk, lam, t, rm, rd, rg[6], rb[3]; };
  __device__ __forceinline__ S1v load_s1(int k) const {
    S1v v;
    v.mk = mask_of(k); v.lam = ws_ld(wsc(k, v2::oLAM)); v.t = ws_ld(wsc(k, v2::oT));
    v.rm = ws_ld(wsc(k, v2::oRM)); v.rd = ws_ld(wsc(k, v2::oRD));
#pragma unroll
    for (int j = 0; j < 6; ++j) v.rg[j] = ws_ld(wsf(k, v2::oRG) + 4 * j);
#pragma unroll
    for (int j = 0; j < 3; ++j) v.rb[j] = ws_ld(wsf(k, v2::oRB) + 4 * j);
    return v;
  }
  // 1/sqrt(x) of a positive pivot: MUFU.RSQ64H seed (rel. error < 2^-20) + one third-order step
  // y (1 + e/2 + 3 e^2/8), e = 1 - x y^2  (error ~ e^3); a non-positive pivot gives 0 like BLASFEO's potrf
  // (BLASFEO potrf_l_mn semantics)
  static __device__ __forceinline__ double rsqrt_pivot(double x) {
#if SRBD_K3_EXACT_RSQRT == 1
    return x > 0.0 ? 1.0 / sqrt(x) : 0.0;   // (experiment: IEEE sqrt + division)
#elif SRBD_K3_EXACT_RSQRT == 2
    return x > 0.0 ? rsqrt(x) : 0.0;        // (experiment: the CUDA library function, 1 ulp)
#endif
    if (kPivot == 1) return x > 0.0 ? rsqrt(x) : 0.0;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
    const double tt = x * y;
    const double e = fma(-tt, y, 1.0);
    const double q = e * fma(0.375, e, 0.5);
    y = fma(y, q, y);
#if SRBD_K3_EXACT_RSQRT == 3
    {  // (experiment: one more Newton step on the residual of the refined value)
      const double t2 = x * y;
      const double e2 = fma(-t2, y, 1.0);
      y = fma(y * 0.5, e2, y);
    }
#endif
    return x > 0.0 ? y : 0.0;
  }
  __device__ void sweep_factor() {
    constexpr bool kT1 = (kTma & 1) != 0;
    const double reg = p.a.reg_prim;
    double* sP = sm + v2::wP;
    double* sPan = sm + v2::wPAN;
    const int r = fr, t = ft, pi = fpi;
    // per-lane fragment offsets: BAbt tile (A operand rows r, B operand rows pi(r)), panel rows (M row m lives in
    // panel row (m < 12 ? 25 + m : m), E row i in row i, the gradient in row 24)
    const int gA = (r >> 2) * v2::kGP + 4 * t + (r & 3);
    const int gB = (pi >> 2) * v2::kGP + 4 * t + (pi & 3);
    const int prA0 = (25 + r) * 4 + t, prA1 = (r < 4 ? 33 + r : 8 + r) * 4 + t, prA2 = (16 + r) * 4 + t;
    const int prB0 = (25 + pi) * 4 + t, prB1 = (pi < 4 ? 33 + pi : 8 + pi) * 4 + t, prB2 = (16 + pi) * 4 + t;
    const int prE0 = r * 4 + t;
    const int oDt = t * 12 + pi;  // Ac[4kt+t][8I+pi] : + 8 I + 48 kt
    const bool dg0 = (r == t), dg1 = (r == 4 + t);
    const double regd0 = dg0 ? reg : 0.0, regd1 = dg1 ? reg : 0.0;
    const int rel0 = ((r >> 2) ? 16 : 0) + (r & 3) + 4 * t;  // R(i, c) of rows 0..7 in the R tile:  + 16 * (c >> 2)
    const int rel1 = 48 + (r & 3) + 4 * t;                   // rows 8..11 (r < 4)
    const double qd3 = dg1 ? cQ[t] : 0.0, qd4 = dg0 ? cQ[4 + t] : 0.0, qd5 = dg1 ? cQ[8 + t] : 0.0;  // diag(Q)
    // start streaming stage N-1 while stage N is handled
    tiles_begin<kT1>();
    prefetch_G<kT1>(N - 1, 0);
    prefetch_Rblk<kT1>(N - 1, 0);
    S1v cur = load_s1(N - 1);
    // ---- stage N: P_N = Q_N + reg I, p_N = rg_N ------------------------------------------------------
    double pk[3];
    {
      const double* rs = gRecL(N) - kLaneOff * lane;  // Q_N: the R tile of the last stage record
      if (lane < 12) {
#pragma unroll
        for (int c = 0; c < 12; ++c) {
          if (c <= lane) {
            double v = __ldg(rs + ((lane >> 2) == 0 ? 0 : ((lane >> 2) == 1 ? 16 : 48)) + 4 * c + (lane & 3));
            if (c == lane) v += reg;
            sP[lane * 12 + c] = v;
            sP[c * 12 + lane] = v;
          }
        }
      }
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) pk[kt] = ws_ld(wsf(N, v2::oRG) + 4 * kt);
      __syncwarp();
      for (int e = lane; e < 144; e += 32) ws(N - 1, v2::oP)[e] = sP[e];  // P_k lives in block k-1 (next to FT_{k-1})
      if (r == 0) {
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) wsf(N, v2::oPV)[4 * kt] = pk[kt];
      }
    }
#if SRBD_K3_UNROLL_FAC
#pragma unroll 2
#else
#pragma unroll (kTeam ? 2 : 1)   // latency mode: the leader warp has the SM (and its instruction cache) to itself
#endif
    for (int k = N - 1; k >= 0; --k) {
      const int b = (N - 1 - k) & 1;
      const bool xr = k > 0;                 // the stage has x rows (rows 12..23)
      tiles_wait<kT1>(b);
      __syncwarp();  // stage k's tiles have landed; every lane is done with the other buffers
      set_bufs(b);
      S1v nxt = cur;
      if (k > 0) {
        tiles_begin<kT1>();
        prefetch_G<kT1>(k - 1, b ^ 1);
        prefetch_Rblk<kT1>(k - 1, b ^ 1);
        nxt = load_s1(k - 1);
        if (k > 1) prefetch_L2(k - 2, false, false, true);
      }
      double* gbuf = sm + (b ? v2::wqx : v2::wQX);
      double* Gbuf = sm + (b ? v2::wSG : v2::wSX);
      if (lane < 24) {
        const double ti = 1.0 / cur.t;
        Gbuf[lane] = (ti * cur.lam) * cur.mk;
        gbuf[lane] = (ti * (cur.rm - cur.lam * cur.rd)) * cur.mk;
      }
      __syncwarp();
      // ---- operand fragments: G (A operand), G rows permuted (B operand), P_{k+1} (B operand) -------------------
      double GF[3][3], GPF[3][3], PPF[2][3];
      {
        const bool okA1 = xr || r < 4, okB1 = xr || pi < 4;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          GF[0][kt] = sG[gA + 16 * kt];
          const double a1 = sG[gA + 2 * v2::kGP + 16 * kt], a2 = sG[gA + 4 * v2::kGP + 16 * kt];
          GF[1][kt] = okA1 ? a1 : 0.0;
          GF[2][kt] = xr ? a2 : 0.0;
          GPF[0][kt] = sG[gB + 16 * kt];
          const double b1 = sG[gB + 2 * v2::kGP + 16 * kt], b2 = sG[gB + 4 * v2::kGP + 16 * kt];
          GPF[1][kt] = okB1 ? b1 : 0.0;
          GPF[2][kt] = xr ? b2 : 0.0;
          PPF[0][kt] = sP[pi * 12 + 4 * kt + t];
          const double p1 = sP[(8 + (pi & 3)) * 12 + 4 * kt + t];
          PPF[1][kt] = pi < 4 ? p1 : 0.0;
        }
      }
      // ---- the 42 sums of D^T Gamma D (two 6x6 blocks, lower triangles) -------------------------------------------
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int e2 = lane + 32 * rr;
        if (e2 < 42) {
          const int leg = e2 >= 21 ? 1 : 0, e = e2 - 21 * leg;
          double acc = 0.0;
#pragma unroll
          for (int g = 0; g < 12; ++g) acc += Gbuf[12 * leg + g] * cW[(12 * leg + g) * v2::kW2 + e];
          sm[v2::wS + e2] = acc;
        }
      }
      // ---- gradient, fragment form: t = P_{k+1} rb + p_{k+1};  g~ = rg + D^T gamma + G t ----------------------------
      double c0[2] = {cur.rg[0], cur.rg[1]}, c1[2] = {cur.rg[2], cur.rg[3]}, c2[2] = {cur.rg[4], cur.rg[5]};
      {
        double d0[2] = {0.0, 0.0}, d1[2] = {0.0, 0.0};
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(d0[0], d0[1], cur.rb[kt], PPF[0][kt], d0[0], d0[1]);
          dmma(d1[0], d1[1], cur.rb[kt], PPF[1][kt], d1[0], d1[1]);
        }
        if (r == 0) {  // P_{k+1} rb is the same for every KKT solve of this iteration
          wsf(k, v2::oPRB)[0] = d0[0]; wsf(k, v2::oPRB)[4] = d0[1]; wsf(k, v2::oPRB)[8] = d1[0];
        }
        const double tk[3] = {d0[0] + pk[0], d0[1] + pk[1], d1[0] + pk[2]};
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
          const double gam = gbuf[4 * kt + t];
#if SRBD_K3_MERGE_HALVES
          // the rows of leg 1 (kt >= 3) only reach columns 6..11 = the second half of tile 0 and the first half of
          // tile 1: one DMMA whose first-half lanes supply Ac[.][8 + pi] and whose second-half lanes Ac[.][pi]
          if (kt < 3) dmma(c0[0], c0[1], gam, cAc[oDt + 48 * kt], c0[0], c0[1]);
          else dmma(c1[0], c0[1], gam, cAc[oDt + 48 * kt + (pi < 4 ? 8 : 0)], c1[0], c0[1]);
#else
          dmma(c0[0], c0[1], gam, cAc[oDt + 48 * kt], c0[0], c0[1]);
          if (kt >= 3) dmma(c1[0], c1[1], gam, cAc[oDt + 48 * kt + 8], c1[0], c1[1]);
#endif
        }
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(c0[0], c0[1], tk[kt], GPF[0][kt], c0[0], c0[1]);
          dmma(c1[0], c1[1], tk[kt], GPF[1][kt], c1[0], c1[1]);
          dmma(c2[0], c2[1], tk[kt], GPF[2][kt], c2[0], c2[1]);
        }
      }
      // ---- AL = G P_{k+1}: fragments ALF[I][0..2] (columns 0..3, 4..7, 8..11) -------------------------------------
      double ALF[3][3];
#pragma unroll
      for (int I = 0; I < 3; ++I) {
        double a0 = 0.0, a1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(a0, a1, GF[I][kt], PPF[0][kt], a0, a1);
          dmma(e0, e1, GF[I][kt], PPF[1][kt], e0, e1);
        }
        ALF[I][0] = a0; ALF[I][1] = a1; ALF[I][2] = e0;
      }
      // ---- M = (H + D^T Gamma D) + AL G^T + reg I: lower 8x8 tiles, MF[I][2J], MF[I][2J+1] ------------------------
      double MF[3][6];
#pragma unroll
      for (int I = 0; I < 3; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J) {
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int kt = 0; kt < 3; ++kt) dmma(a0, a1, ALF[I][kt], GPF[J][kt], a0, a1);
          MF[I][2 * J] = a0; MF[I][2 * J + 1] = a1;
        }
      __syncwarp();  // wS visible
      {
        // u block (rows < 12): R_k (lower) + D^T Gamma D, fragments (0,0) (0,1) (1,0) (1,1) (1,2)
        double hf[5];
#pragma unroll
        for (int f = 0; f < 5; ++f) {
          const int I = f < 2 ? 0 : 1, pp = f < 2 ? f : f - 2;
          const bool valid = (8 * I + r < 12) && (4 * pp + t <= 8 * I + r);
          double v = 0.0;
          if (valid) v = sR[(I == 0 ? rel0 : rel1) + 16 * pp];
          if (offS[f] >= 0) v += sm[v2::wS + offS[f]];
          hf[f] = v;
        }
        MF[0][0] = (hf[0] + MF[0][0]) + regd0;
        MF[0][1] = (hf[1] + MF[0][1]) + regd1;
        MF[1][0] = hf[2] + MF[1][0];
        MF[1][1] = hf[3] + MF[1][1];
        MF[1][2] = (hf[4] + MF[1][2]) + regd0;
        // x block: diagonal of Q (rows 12..23)
        MF[1][3] = ((xr ? qd3 : 0.0) + MF[1][3]) + regd1;
        MF[2][4] = ((xr ? qd4 : 0.0) + MF[2][4]) + regd0;
        MF[2][5] = ((xr ? qd5 : 0.0) + MF[2][5]) + regd1;
      }
      // ---- blocked Cholesky of the 12 u columns: three 4-column panels ------------------------------------------
#pragma unroll
      for (int pp = 0; pp < 3; ++pp) {
        double* pan = sPan + pp * v2::kPan;
        if (pp < 2) pan[prA0] = MF[0][pp];
        pan[prA1] = MF[1][pp];
        pan[prA2] = MF[2][pp];
        // E rows: the identity block of rows 4pp..4pp+3 -> the substitution leaves L_pp^-T there (the diagonal
        // blocks of the blocked triangular solves of the vector sweeps); the other E rows are zero
        pan[prE0] = pp == 0 ? (dg0 ? 1.0 : 0.0) : ((pp == 1 && dg1) ? 1.0 : 0.0);
        if (r < 4) pan[(8 + r) * 4 + t] = (pp == 2 && dg0) ? 1.0 : 0.0;  // E rows 8..11
        if (r == 0) pan[96 + t] = pp == 0 ? c0[0] : (pp == 1 ? c0[1] : c1[0]);  // gradient row: g~[4pp + t]
        __syncwarp();
        // every lane factors the 4x4 diagonal block (rows 25+4pp..) redundantly ...
        const double* dgb = pan + (25 + 4 * pp) * 4;
        const double a00 = dgb[0];
        const double2 a1x = *reinterpret_cast<const double2*>(dgb + 4);
        const double2 a2x = *reinterpret_cast<const double2*>(dgb + 8);
        const double a22 = dgb[10];
        const double2 a3x = *reinterpret_cast<const double2*>(dgb + 12);
        const double2 a3y = *reinterpret_cast<const double2*>(dgb + 14);
        // ... and substitutes its own row: M rows 4pp..23, the gradient (lane 24), E rows 0..4pp+3
        const int prow = lane < 4 * pp ? 4 + lane : (lane < 12 ? 25 + lane : (lane < 25 ? lane : lane - 25));
        double2* own = reinterpret_cast<double2*>(pan + prow * 4);
        const double2 x01 = own[0], x23 = own[1];
        const double i0 = rsqrt_pivot(a00);
        const double L10 = a1x.x * i0, L20 = a2x.x * i0, L30 = a3x.x * i0;
        const double i1 = rsqrt_pivot(fma(-L10, L10, a1x.y));
        const double L21 = fma(-L20, L10, a2x.y) * i1, L31 = fma(-L30, L10, a3x.y) * i1;
        const double i2 = rsqrt_pivot(fma(-L21, L21, fma(-L20, L20, a22)));
        const double L32 = fma(-L31, L21, fma(-L30, L20, a3y.x)) * i2;
        const double i3 = rsqrt_pivot(fma(-L32, L32, fma(-L31, L31, fma(-L30, L30, a3y.y))));
        const double l0 = x01.x * i0;
        const double l1 = fma(-l0, L10, x01.y) * i1;
        const double l2 = fma(-l1, L21, fma(-l0, L20, x23.x)) * i2;
        const double l3 = fma(-l2, L32, fma(-l1, L31, fma(-l0, L30, x23.y))) * i3;
        own[0] = make_double2(l0, l1);
        own[1] = make_double2(l2, l3);
        __syncwarp();
        // trailing updates M -= Lp Lp^T and g~ -= Lp lv_p on the tensor cores (A = -L panel fragment / -lv_p in
        // row 0, B = permuted fragment of the panel)
        const double nlv = -pan[96 + t];
        double dum;
        // (no branch on the stage type: at stage 0 the rows 12..23 are zero and the extra DMMAs produce zeros; the
        // updates that feed the NEXT panel come first, the rest trails behind the next panel's substitution)
        if (pp == 0) {
          const double nA0 = -pan[prA0], nA1 = -pan[prA1], nA2 = -pan[prA2];
          const double B0 = pan[prB0], B1 = pan[prB1], B2 = pan[prB2];
          dmma(dum, MF[0][1], nA0, B0, MF[0][0], MF[0][1]);
          dmma(dum, MF[1][1], nA1, B0, MF[1][0], MF[1][1]);
          dmma(dum, MF[2][1], nA2, B0, MF[2][0], MF[2][1]);
          dmma(dum, c0[1], nlv, B0, c0[0], c0[1]);
          dmma(MF[1][2], MF[1][3], nA1, B1, MF[1][2], MF[1][3]);
          dmma(MF[2][2], MF[2][3], nA2, B1, MF[2][2], MF[2][3]);
          dmma(c1[0], c1[1], nlv, B1, c1[0], c1[1]);
          dmma(MF[2][4], MF[2][5], nA2, B2, MF[2][4], MF[2][5]);
          dmma(c2[0], c2[1], nlv, B2, c2[0], c2[1]);
        } else if (pp == 1) {
          const double nA1 = -pan[prA1], nA2 = -pan[prA2];
          const double B1 = pan[prB1], B2 = pan[prB2];
          dmma(MF[1][2], MF[1][3], nA1, B1, MF[1][2], MF[1][3]);
          dmma(MF[2][2], MF[2][3], nA2, B1, MF[2][2], MF[2][3]);
          dmma(c1[0], c1[1], nlv, B1, c1[0], c1[1]);
          dmma(MF[2][4], MF[2][5], nA2, B2, MF[2][4], MF[2][5]);
          dmma(c2[0], c2[1], nlv, B2, c2[0], c2[1]);
        } else {
          const double nA1 = -pan[prA1], nA2 = -pan[prA2];
          const double B1 = pan[prB1], B2 = pan[prB2];
          dmma(dum, MF[1][3], nA1, B1, MF[1][2], MF[1][3]);
          dmma(dum, MF[2][3], nA2, B1, MF[2][2], MF[2][3]);
          dmma(MF[2][4], MF[2][5], nA2, B2, MF[2][4], MF[2][5]);
          dmma(dum, c1[1], nlv, B1, c1[0], c1[1]);
          dmma(c2[0], c2[1], nlv, B2, c2[0], c2[1]);
        }
      }
      // ---- outputs: P_k (both triangles) and p_k for the next stage and the vector sweeps --------------------------
      if (xr) {
        // fragments (1,3): rows 12..15 x cols 12..15 (r >= 4); (2,3): rows 16..23 x cols 12..15;
        // (2,4): x cols 16..19; (2,5): x cols 20..23
        if (r >= 4 && t <= r - 4) { sP[(r - 4) * 12 + t] = MF[1][3]; sP[t * 12 + (r - 4)] = MF[1][3]; }
        sP[(4 + r) * 12 + t] = MF[2][3]; sP[t * 12 + (4 + r)] = MF[2][3];
        if (t <= r) { sP[(4 + r) * 12 + 4 + t] = MF[2][4]; sP[(4 + t) * 12 + 4 + r] = MF[2][4]; }
        if (4 + t <= r) { sP[(4 + r) * 12 + 8 + t] = MF[2][5]; sP[(8 + t) * 12 + 4 + r] = MF[2][5]; }
        pk[0] = c1[1]; pk[1] = c2[0]; pk[2] = c2[1];   // p_k = g~_x after the three panels
        if (r == 0) {
          wsf(k, v2::oPV)[0] = pk[0]; wsf(k, v2::oPV)[4] = pk[1]; wsf(k, v2::oPV)[8] = pk[2];
        }
        __syncwarp();
        for (int e = lane; e < 72; e += 32)
          reinterpret_cast<double2*>(ws(k - 1, v2::oP))[e] = reinterpret_cast<const double2*>(sP)[e];
      }
      // factor panels -> workspace.  Panel pp, rows 0..11: row i in [4pp, 4pp+4) = L_pp^-T (E rows), row i >= 4pp+4 =
      // L[i][4pp..4pp+3] (panel rows 25+i); rows 12..23 = Ls, row 24 = lv
#pragma unroll
      for (int pp = 0; pp < 3; ++pp) {
        const double2* src = reinterpret_cast<const double2*>(sPan + pp * v2::kPan);
        double2* dst = reinterpret_cast<double2*>(ws(k, v2::oFT) + pp * v2::kPanF);
        const int row = lane >> 1;
        dst[lane] = src[(row < 12 && row >= 4 * pp + 4) ? lane + 50 : lane];
        if (lane < 18) dst[32 + lane] = src[32 + lane];
      }
      cur = nxt;
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------------
  // Vector sweeps in FRAGMENT FORM (scripts/proto_dmma_vec.py is the lane-level emulation of these bodies).
  // A length-12 vector is three registers ("k-tiles") valid in lanes 0..3: lane t of k-tile kt holds v[4 kt + t],
  // i.e. row 0 of a DMMA A operand (the other rows replicate it).  y = A x is 2 x 3 DMMA with the rows of A as
  // pi-permuted B fragments read from the cp.async-staged shared-memory tiles (fragments gathered straight from
  // global memory into registers were slower: DESIGN.md section 5).  The accumulator pair (c0, c1) of output tile I
  // is k-tiles 2I, 2I+1 of y, so a whole stage of the recursion (three to five chained gemvs) runs in registers;
  // the only shared-memory round trip per stage is between the fragment form and the row-per-lane form of the 24
  // constraint rows.
  // ------------------------------------------------------------------------------------------------
  // S4: vector-only backward sweep (gradient recursion with the stored factors).  mode 1: centering
  // correction, mode 2: centering only; sm_ = sigma*mu (clamped by the caller)
  struct S4v { double mk, dt, dlam, lam, t, rd, rg[6], prb[3]; };
  __device__ __forceinline__ S4v load_s4(int k) const {
    S4v v;
    v.mk = mask_of(k); v.dt = ws_ld(wsc(k, v2::oDT));
    v.dlam = ws_ld(wsc(k, v2::oDLAM)); v.lam = ws_ld(wsc(k, v2::oLAM)); v.t = ws_ld(wsc(k, v2::oT));
    v.rd = ws_ld(wsc(k, v2::oRD));
#pragma unroll
    for (int j = 0; j < 6; ++j) v.rg[j] = ws_ld(wsf(k, v2::oRG) + 4 * j);
#pragma unroll
    for (int j = 0; j < 3; ++j) v.prb[j] = ws_ld(wsf(k, v2::oPRB) + 4 * j);
    return v;
  }
  __device__ __forceinline__ void sweep_backvec(int mode, double sm_) {
    constexpr bool kT4 = (kTma & 2) != 0;
    const int r = fr, t = ft, pi = fpi;
    const int oG = (pi >> 2) * v2::kGP + 4 * t + (pi & 3);            // G[8I+pi][4kt+t]      : + 2 kGP I + 16 kt
    // 4x4 blocks of the factor panels as B fragments: row index t / column pi, and row pi / column t
    const int bA = 144 + 4 * t + (pi & 3), bB = 144 + 4 * (pi & 3) + t;   // + kPanF * (column panel) + 16 * (row block)
    const int oLs = 144 + (12 + pi) * 4 + t;                          // Ls[8I+pi][4kt+t]     : + 32 I + kPanF kt
    const int oDt = t * 12 + pi;                                      // Ac[4kt+t][8I+pi]     : + 8 I + 48 kt
    double pk[3];
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) pk[kt] = ws_ld(wsf(N, v2::oRG) + 4 * kt);
    if (r == 0) {
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) wsf(N, v2::oPV)[4 * kt] = pk[kt];
    }
    tiles_begin<kT4>();
    prefetch_G<kT4>(N - 1, 0);
    prefetch_F<kT4>(N - 1, 0, false);
    S4v cur = load_s4(N - 1);
#if SRBD_K3_UNROLL_BWD
#pragma unroll 2
#else
#pragma unroll (kTeam ? 2 : 1)   // latency mode: the leader warp has the SM (and its instruction cache) to itself
#endif
    for (int k = N - 1; k >= 0; --k) {
      const int b = (N - 1 - k) & 1;
      tiles_wait<kT4>(b);
      __syncwarp();  // stage k's tiles have landed; every lane is done with the other buffers
      set_bufs(b);
      S4v nxt = cur;
      if (k > 0) {
        tiles_begin<kT4>();
        prefetch_G<kT4>(k - 1, b ^ 1);
        prefetch_F<kT4>(k - 1, b ^ 1, false);
        nxt = load_s4(k - 1);
        if (k > 1) prefetch_L2(k - 2, false, true, false);
      }
      double* gbuf = sm + (b ? v2::wqx : v2::wQX);
      // ---- row-per-lane: res_m of this solve and gamma ------------------------------------------------------------
      {
        // res_m of the iterate (BACKUP_RES_M) = lam * t on the active rows: recomputed from the iterate the residual sweep
        // stored (bit-identical to what it put into res_m) instead of a stored copy
        double rm = (cur.lam * cur.t) * cur.mk;
        if (mode == 1) rm += cur.dt * cur.dlam;
        rm = (rm - sm_) * cur.mk;
        const double ti = 1.0 / cur.t;
        if (lane < 24) {
          wsc(k, v2::oRM)[0] = rm;
          gbuf[lane] = (ti * (rm - cur.lam * cur.rd)) * cur.mk;
        }
      }
      __syncwarp();
      // ---- t = P_{k+1} rb + p_{k+1}  (P rb was stored by the factorization sweep) ------------------------------------
      double tk[3];
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) tk[kt] = cur.prb[kt] + pk[kt];
      // ---- g~ = rg + D^T gamma + G t  (tiles of rows 0..7, 8..15, 16..23) ------------------------------------------
      double c0[2] = {cur.rg[0], cur.rg[1]}, c1[2] = {cur.rg[2], cur.rg[3]}, c2[2] = {cur.rg[4], cur.rg[5]};
#pragma unroll
      for (int kt = 0; kt < 6; ++kt) {
        const double gam = gbuf[4 * kt + t];
#if SRBD_K3_MERGE_HALVES
        // the rows of leg 1 (kt >= 3) only reach columns 6..11 = the second half of tile 0 and the first half of
        // tile 1: one DMMA whose first-half lanes supply Ac[.][8 + pi] and whose second-half lanes Ac[.][pi]
        if (kt < 3) dmma(c0[0], c0[1], gam, cAc[oDt + 48 * kt], c0[0], c0[1]);
        else dmma(c1[0], c0[1], gam, cAc[oDt + 48 * kt + (pi < 4 ? 8 : 0)], c1[0], c0[1]);
#else
        dmma(c0[0], c0[1], gam, cAc[oDt + 48 * kt], c0[0], c0[1]);
        if (kt >= 3) dmma(c1[0], c1[1], gam, cAc[oDt + 48 * kt + 8], c1[0], c1[1]);
#endif
      }
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        dmma(c0[0], c0[1], tk[kt], sG[oG + 16 * kt], c0[0], c0[1]);
        dmma(c1[0], c1[1], tk[kt], sG[oG + 2 * v2::kGP + 16 * kt], c1[0], c1[1]);
        dmma(c2[0], c2[1], tk[kt], sG[oG + 4 * v2::kGP + 16 * kt], c2[0], c2[1]);
      }
#if SRBD_K3_DUMMY_DMMA
      {  // EXPERIMENT: marginal cost of one more gemv-form DMMA group (9 DMMA on operands that are already loaded)
        double z0[2] = {0.0, 0.0}, z1[2] = {0.0, 0.0}, z2[2] = {0.0, 0.0};
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(z0[0], z0[1], tk[kt], sG[oG + 16 * kt], z0[0], z0[1]);
          dmma(z1[0], z1[1], tk[kt], sG[oG + 2 * v2::kGP + 16 * kt], z1[0], z1[1]);
          dmma(z2[0], z2[1], tk[kt], sG[oG + 4 * v2::kGP + 16 * kt], z2[0], z2[1]);
        }
        if (z0[0] + z1[1] + z2[0] == 123.456789) ws(k, v2::oLV)[0] = z0[1] + z1[0] + z2[1];
      }
#endif
      // ---- lv = L^-1 g~_u: blocked forward substitution (4x4 diagonal blocks by their inverses L_pp^-1, which the
      // factorization left in the E rows; same operation order as trsv up to the blocks) ----------------------------------
      double lv0, lv1, lv2, junk;
      {
        const int P1 = v2::kPanF, P2 = 2 * v2::kPanF;
        double g1, g2;
        dmma(lv0, junk, c0[0], sF[bA], 0.0, 0.0);                       // lv0 = X0 g0
#if SRBD_K3_MERGE_HALVES
        // (one DMMA for both products of -lv0: the lanes of the first accumulator half supply L10, the others L20)
        dmma(g1, g2, -lv0, sF[bB + 16 + (pi < 4 ? 0 : 16)], c0[1], c1[0]);   // g1 - L10 lv0 | g2 - L20 lv0
#else
        dmma(g1, junk, -lv0, sF[bB + 16], c0[1], 0.0);                  // g1 - L10 lv0
        dmma(g2, junk, -lv0, sF[bB + 32], c1[0], 0.0);                  // g2 - L20 lv0
#endif
        dmma(lv1, junk, g1, sF[bA + P1 + 16], 0.0, 0.0);                // lv1 = X1 (.)
        dmma(g2, junk, -lv1, sF[bB + P1 + 32], g2, 0.0);                // ... - L21 lv1
        dmma(lv2, junk, g2, sF[bA + P2 + 32], 0.0, 0.0);                // lv2 = X2 (.)
      }
      if (r == 0) {  // row 24 of the factor panels
        double* lvd = ws(k, v2::oFT) + 96 + t;
        lvd[0] = lv0; lvd[v2::kPanF] = lv1; lvd[2 * v2::kPanF] = lv2;
      }
      // ---- p = g~_x - Ls lv  (g~_x = second half of tile 1 and tile 2: pure register renaming) -----------------------
      if (k > 0) {
        double p0[2] = {c1[1], c2[0]}, p1[2] = {c2[1], 0.0};
        const double nl[3] = {-lv0, -lv1, -lv2};
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(p0[0], p0[1], nl[kt], sF[oLs + v2::kPanF * kt], p0[0], p0[1]);
          dmma(p1[0], p1[1], nl[kt], sF[oLs + 32 + v2::kPanF * kt], p1[0], p1[1]);
        }
        pk[0] = p0[0]; pk[1] = p0[1]; pk[2] = p1[0];
        if (r == 0) {
          wsf(k, v2::oPV)[0] = pk[0]; wsf(k, v2::oPV)[4] = pk[1]; wsf(k, v2::oPV)[8] = pk[2];
        }
      }
      cur = nxt;
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------------
  // S2/S5: forward rollout fused with dt / dlam, the step length and the three sums that give
  // mu_aff(alpha) = sum (lam + alpha dlam)(t + alpha dt) / nc for any alpha.  fin: this may be the last KKT
  // solve of the iteration: also produce dpi and store dz (the predictor needs neither).
  // ------------------------------------------------------------------------------------------------
  double mu_s0, mu_s1, mu_s2;  // sum lam t, sum (lam dt + t dlam), sum dlam dt of the last forward sweep
  struct S2v { double mk, t, lam, rd, rm, rb[3], pv[3]; };
  __device__ __forceinline__ S2v load_s2(int k) const {
    S2v v;
    v.mk = mask_of(k); v.t = ws_ld(wsc(k, v2::oT)); v.lam = ws_ld(wsc(k, v2::oLAM));
    v.rd = ws_ld(wsc(k, v2::oRD)); v.rm = ws_ld(wsc(k, v2::oRM));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      v.rb[j] = ws_ld(wsf(k, v2::oRB) + 4 * j);
      v.pv[j] = ws_ld(wsf(k + 1, v2::oPV) + 4 * j);
    }
    return v;
  }
  __device__ __forceinline__ void sweep_forward(bool fin, double& ap, double& ad) {
    constexpr bool kT2 = (kTma & 4) != 0;
    const int r = fr, t = ft, pi = fpi;
    const int oLT = 144 + (pi >> 2) * v2::kPanF + 48 + 4 * t + (pi & 3);  // Ls[4kt+t][8I+pi]   : + 2 kPanF I + 16 kt
    // 4x4 blocks of the factor panels as B fragments: row index t / column pi, and row pi / column t
    const int bA = 144 + 4 * t + (pi & 3), bB = 144 + 4 * (pi & 3) + t;   // + kPanF * (column panel) + 16 * (row block)
    const int oGT = 4 * pi + t;                                           // G[4kt+t][8I+pi]    : + 32 I + kGP kt
    const int oPP = pi * 12 + t;                                          // P[8I+pi][4kt+t]    : + 96 I + 4 kt
    const int oLV = 144 + 96 + t;                                         // lv[4kt+t]          : + kPanF kt
    const int lc = lane < 24 ? lane : 0;
    const int j0 = lc < 12 ? 0 : 6;
    double acr[6];  // this lane's constraint row (6 nonzeros)
#pragma unroll
    for (int j = 0; j < 6; ++j) acr[j] = cAc[lc * 12 + j0 + j];
    double np_ = 1.0, dp_ = 1.0, nd_ = 1.0, dd_ = 1.0;  // step lengths as fractions num/den (no division per row)
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    double xk[3] = {0.0, 0.0, 0.0};
    tiles_begin<kT2>();
    prefetch_G<kT2>(0, 0);
    prefetch_F<kT2>(0, 0, fin);
    S2v cur = load_s2(0);
#if SRBD_K3_UNROLL_FWD
#pragma unroll 2
#else
#pragma unroll (kTeam ? 2 : 1)   // latency mode: the leader warp has the SM (and its instruction cache) to itself
#endif
    for (int k = 0; k < N; ++k) {
      const int b = k & 1;
      tiles_wait<kT2>(b);
      __syncwarp();
      set_bufs(b);
      S2v nxt = cur;
      if (k + 1 < N) {
        tiles_begin<kT2>();
        prefetch_G<kT2>(k + 1, b ^ 1);
        prefetch_F<kT2>(k + 1, b ^ 1, fin);
        nxt = load_s2(k + 1);
        if (k + 2 < N) prefetch_L2(k + 2, fin, true, false);
      }
      double* ubuf = sm + (b ? v2::wSX : v2::wSG);
      // ---- x+ = G^T [u; x] + rb: the x part first (it does not wait for u);  t = Ls^T x + lv ----------------------------
      double cx0[2] = {cur.rb[0], cur.rb[1]}, cx1[2] = {cur.rb[2], 0.0};
      double t0[2] = {sF[oLV], sF[oLV + v2::kPanF]}, t1[2] = {sF[oLV + 2 * v2::kPanF], 0.0};
      {  // (stage 0: x = 0 and the rows 12.. of its records are finite, so no branch on the stage type)
#pragma unroll
#if SRBD_K3_MERGE_HALVES
        // rows 8..11 of t and of x+ use half an output tile each and multiply the same x: ONE DMMA whose first-half
        // lanes (pi < 4) supply Ls^T and whose second-half lanes supply G^T (same per-element arithmetic, 3 DMMA less)
        const int om = pi < 4 ? (b ? v2::wF1 : v2::wF0) + oLT + 2 * v2::kPanF
                              : (b ? v2::wG1 : v2::wG0) + oGT + 16 + 3 * v2::kGP;
        const int os = pi < 4 ? 16 : v2::kGP;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(t0[0], t0[1], xk[kt], sF[oLT + 16 * kt], t0[0], t0[1]);
          dmma(cx0[0], cx0[1], xk[kt], sG[oGT + v2::kGP * (3 + kt)], cx0[0], cx0[1]);
          dmma(t1[0], cx1[0], xk[kt], sm[om + os * kt], t1[0], cx1[0]);
        }
#else
        for (int kt = 0; kt < 3; ++kt) {
          dmma(t0[0], t0[1], xk[kt], sF[oLT + 16 * kt], t0[0], t0[1]);
          dmma(t1[0], t1[1], xk[kt], sF[oLT + 2 * v2::kPanF + 16 * kt], t1[0], t1[1]);
          dmma(cx0[0], cx0[1], xk[kt], sG[oGT + v2::kGP * (3 + kt)], cx0[0], cx0[1]);
          dmma(cx1[0], cx1[1], xk[kt], sG[oGT + 32 + v2::kGP * (3 + kt)], cx1[0], cx1[1]);
        }
#endif
      }
      // ---- u = -L^-T t: blocked back substitution (diagonal blocks by L_pp^-T) ---------------------------------------
      double uk[3];
      {
        const int P1 = v2::kPanF, P2 = 2 * v2::kPanF;
        double w0, w1, junk;
        dmma(uk[2], junk, -t1[0], sF[bB + P2 + 32], 0.0, 0.0);           // u2 = X2^T (-t2)
#if SRBD_K3_MERGE_HALVES
        dmma(w1, w0, -uk[2], sF[bA + 32 + (pi < 4 ? P1 : 0)], -t0[1], -t0[0]);   // -t1 - L21^T u2 | -t0 - L20^T u2
#else
        dmma(w1, junk, -uk[2], sF[bA + P1 + 32], -t0[1], 0.0);           // -t1 - L21^T u2
        dmma(w0, junk, -uk[2], sF[bA + 32], -t0[0], 0.0);                // -t0 - L20^T u2
#endif
        dmma(uk[1], junk, w1, sF[bB + P1 + 16], 0.0, 0.0);               // u1 = X1^T (.)
        dmma(w0, junk, -uk[1], sF[bA + 16], w0, 0.0);                    // ... - L10^T u1
        dmma(uk[0], junk, w0, sF[bB], 0.0, 0.0);                         // u0 = X0^T (.)
      }
      if (r == 0) {  // row-per-lane consumers read u from shared memory
        ubuf[t] = uk[0]; ubuf[4 + t] = uk[1]; ubuf[8 + t] = uk[2];
      }
      // ---- x+ += B^T u --------------------------------------------------------------------------------------------------
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        dmma(cx0[0], cx0[1], uk[kt], sG[oGT + v2::kGP * kt], cx0[0], cx0[1]);
        dmma(cx1[0], cx1[1], uk[kt], sG[oGT + 32 + v2::kGP * kt], cx1[0], cx1[1]);
      }
      if (fin && r == 0) {
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          wsf(k, v2::oDZ)[4 * kt] = uk[kt];
          if (k > 0) wsf(k, v2::oDZ)[12 + 4 * kt] = xk[kt];
        }
      }
      xk[0] = cx0[0]; xk[1] = cx0[1]; xk[2] = cx1[0];
      // ---- dpi = P_{k+1} x+ + p_{k+1} ----------------------------------------------------------------------------------
      if (fin) {
        double d0[2] = {cur.pv[0], cur.pv[1]}, d1[2] = {cur.pv[2], 0.0};
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(d0[0], d0[1], xk[kt], sF[oPP + 4 * kt], d0[0], d0[1]);
          dmma(d1[0], d1[1], xk[kt], sF[oPP + 96 + 4 * kt], d1[0], d1[1]);
        }
        if (r == 0) {
          wsf(k, v2::oDPI)[0] = d0[0]; wsf(k, v2::oDPI)[4] = d0[1]; wsf(k, v2::oDPI)[8] = d1[0];
        }
      }
      __syncwarp();
      // ---- constraint rows (row-per-lane): v = D du, dt, dlam, step lengths, mu_aff sums ------------------------------
      {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) v += acr[j] * ubuf[j0 + j];
        const double dt = (v - cur.rd) * cur.mk;
        // (masked rows: a zero numerator would send these lanes through the division's slow path every stage)
        const double num = cur.mk != 0.0 ? -(cur.lam * dt + cur.rm) : 1.0;
        const double dl = (num / cur.t) * cur.mk;
        if (lane < 24) {
          wsc(k, v2::oDT)[0] = dt;
          wsc(k, v2::oDLAM)[0] = dl;
          s0 += cur.lam * cur.t;
          s1 += cur.lam * dt + cur.t * dl;
          s2 += dl * dt;
        }
        // alpha = min(1, min -t/dt): keep the minimizer as a fraction, compare by cross-multiplication
        if (dt < 0.0 && cur.t * dp_ < np_ * (0.0 - dt)) { np_ = cur.t; dp_ = 0.0 - dt; }
        if (dl < 0.0 && cur.lam * dd_ < nd_ * (0.0 - dl)) { nd_ = cur.lam; dd_ = 0.0 - dl; }
      }
      cur = nxt;
    }
    if (fin && r == 0) {  // x_N
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) wsf(N, v2::oDZ)[4 * kt] = xk[kt];
    }
    ap = warp_min(np_ / dp_);
    ad = warp_min(nd_ / dd_);
    mu_s0 = warp_sum(s0); mu_s1 = warp_sum(s1); mu_s2 = warp_sum(s2);
    __syncwarp();
  }

  __device__ __forceinline__ double mu_aff(double alpha, int nc_mask) const {
    return ((mu_s2 * alpha + mu_s1) * alpha + mu_s0) / (double)nc_mask;
  }

  // ------------------------------------------------------------------------------------------------
  // S6: d_update_var_qp fused with d_ocp_qp_res_compute + inf norms.  With do_update the new iterate
  // (z,t) += sp*(dz,dt), (pi,lam) += sd*(dpi,dlam) is formed on the fly from the prefetched vectors (and
  // stored), so the update costs no extra pass; res_m is backed up (BACKUP_RES_M) in the same pass.
  // ------------------------------------------------------------------------------------------------
  struct S6v { double z, pi, lam, t, xn, lo, mk; };
  // raw loads of a stage (issued one stage ahead: NO arithmetic on them here, or the warp would wait for the loads at
  // the prefetch point) ...
  struct S6raw { double z, pi, lam, t, xn, lo, mk, dz, dpi, dlam, dt, dxn, ppi[3], pdpi[3]; };
  __device__ __forceinline__ S6raw load_s6(int k, bool do_update) const {
    S6raw v;
    v.z = ws_ld(zc(k, zs));
#pragma unroll
    for (int j = 0; j < 3; ++j) { v.ppi[j] = 0.0; v.pdpi[j] = 0.0; }
    if (kTeam && k > 0) {   // a strided sweep does not carry pi_{k-1} over from its previous trip
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        v.ppi[j] = ws_ld(pif(k - 1, zs) + 4 * j);
        if (do_update) v.pdpi[j] = ws_ld(wsf(k - 1, v2::oDPI) + 4 * j);
      }
    }
    v.pi = 0.0; v.lam = 0.0; v.t = 1.0; v.xn = 0.0; v.lo = 0.0; v.mk = 0.0;
    v.dz = 0.0; v.dpi = 0.0; v.dlam = 0.0; v.dt = 0.0; v.dxn = 0.0;
    const int xo = (k + 1 < N ? 12 : 0);
    if (k < N) {
      v.pi = ws_ld(pic(k, zs)); v.lam = ws_ld(wsc(k, v2::oLAM)); v.t = ws_ld(wsc(k, v2::oT));
      v.xn = ws_ld(zc(k + 1, zs) + xo); v.lo = __ldg(gDL(k)); v.mk = mask_of(k);
    }
    if (do_update) {
      v.dz = ws_ld(wsc(k, v2::oDZ));
      if (k < N) {
        v.dpi = ws_ld(wsc(k, v2::oDPI));
        v.dxn = ws_ld(wsc(k + 1, v2::oDZ) + xo);
        v.dt = ws_ld(wsc(k, v2::oDT));
        v.dlam = ws_ld(wsc(k, v2::oDLAM));
      }
    }
    return v;
  }
  // ... and the variable update d_update_var_qp applied to them when the stage is processed
  __device__ __forceinline__ S6v updated(const S6raw& w, bool do_update, double sp, double sd) const {
    S6v v;
    v.z = w.z; v.pi = w.pi; v.lam = w.lam; v.t = w.t; v.xn = w.xn; v.lo = w.lo; v.mk = w.mk;
    if (do_update) {
      v.z += sp * w.dz;
      v.pi += sd * w.dpi;
      v.xn += sp * w.dxn;
      v.t += sp * w.dt;
      v.lam += sd * w.dlam;
      if (p.a.t_lam_min == 2 && v.mk != 0.0) {
        v.t = v.t < p.a.t_min ? p.a.t_min : v.t;
        v.lam = v.lam < p.a.lam_min ? p.a.lam_min : v.lam;
      }
    }
    return v;
  }
  __device__ void residuals(double res[4], double& mu, double& obj, int nc_mask, bool do_update, double sp, double sd) {
    double acc[6];
    residual_sweep(0, 1, do_update, sp, sd, acc);
    obj = (kExp && p.stat) ? warp_sum(acc[5]) : 0.0;
    const double ng_ = acc[0], nb_ = acc[1], nd_ = acc[2], nm_ = acc[3], smu = acc[4];
    const double flag = warp_sum((ng_ != ng_ || nb_ != nb_ || nd_ != nd_ || nm_ != nm_) ? 1.0 : 0.0);
    res[0] = warp_max(ng_ == ng_ ? ng_ : 0.0);
    res[1] = warp_max(nb_ == nb_ ? nb_ : 0.0);
    res[2] = warp_max(nd_ == nd_ ? nd_ : 0.0);
    res[3] = warp_max(nm_ == nm_ ? nm_ : 0.0);
    if (flag > 0.0) res[0] = res[0] + __longlong_as_double(0x7ff8000000000000LL);
    mu = warp_sum(smu) / (double)nc_mask;
    __syncwarp();
  }
  // Team mode: every warp runs the residual sweep (with the fused variable update) over its stages k = wid, wid + kTeam, ...:
  // it reads (z, pi) of its neighbours from slot zs and writes its own updated ones to slot zs ^ 1 (t, lam, the steps and
  // the residuals are private to a stage).  The partial norms meet in shared memory; the leader re-sums res_m in stage order
  // for mu (bit-identical to the one-warp sweep).  Every warp returns the same res[].
  __device__ void residuals_team(double res[4], double& mu, double& obj, int nc_mask, bool do_update, double sp, double sd) {
    double acc[6];
    residual_sweep(wid, kTeam, do_update, sp, sd, acc);
    if (do_update) zs ^= 1;

    const double ng_ = acc[0], nb_ = acc[1], nd_ = acc[2], nm_ = acc[3];
    const double flag = warp_sum((ng_ != ng_ || nb_ != nb_ || nd_ != nd_ || nm_ != nm_) ? 1.0 : 0.0);
    const double r0 = warp_max(ng_ == ng_ ? ng_ : 0.0), r1 = warp_max(nb_ == nb_ ? nb_ : 0.0);
    const double r2 = warp_max(nd_ == nd_ ? nd_ : 0.0), r3 = warp_max(nm_ == nm_ ? nm_ : 0.0);
    const double ob = (kExp && p.stat) ? warp_sum(acc[5]) : 0.0;
    if (lane == 0) {
      double* o = cred + 6 * wid;
      o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = flag; o[5] = ob;
    }
    __threadfence_block();
    __syncthreads();
    double f = 0.0;
    res[0] = res[1] = res[2] = res[3] = 0.0;
    obj = 0.0;
    for (int w = 0; w < kTeam; ++w) {
      const double* o = cred + 6 * w;
#pragma unroll
      for (int i = 0; i < 4; ++i) res[i] = fmax(res[i], o[i]);
      f += o[4];
      obj += o[5];
    }
    if (f > 0.0) res[0] = res[0] + __longlong_as_double(0x7ff8000000000000LL);
    // mu: the sum of res_m in the order of the one-warp sweep (per row over the stages, then across the rows)
    double smu = 0.0;
    if (wid == 0 && lane < 24) {
#pragma unroll 8
      for (int k = 0; k < N; ++k) smu += ws_ld(wsc(k, v2::oRM));
    }
    mu = warp_sum(smu) / (double)nc_mask;
    __syncwarp();
  }
  // the residual sweep over the stages k0, k0 + kstep, ... <= N; acc: per-lane partial inf-norms (stat, eq, ineq, comp) and
  // the per-lane sum of res_m; with the statistics table also the per-lane part of the objective 1/2 z'Hz + g'z
  __device__ void residual_sweep(int k0, int kstep, bool do_update, double sp, double sd, double acc[6]) {
    const bool want_obj = kExp && p.stat != nullptr;
    double ob = 0.0;
    constexpr bool kT6 = (kTma & 8) != 0;
    const int r = fr, t = ft, pi = fpi;
    const int gB = (pi >> 2) * v2::kGP + 4 * t + (pi & 3);   // G[8I+pi][4kt+t]   : + 2 kGP I + 16 kt
    const int oGT = 4 * pi + t;                              // G[4kt+t][8I+pi]   : + 32 I + kGP kt
    const int oDt = t * 12 + pi;                             // Ac[4kt+t][8I+pi]  : + 8 I + 48 kt
    // the symmetric R block (Q_N at stage N) as B fragments from the lower-triangle tile: element (8I+pi, 4kt+t)
    int oRs[2][3];
#pragma unroll
    for (int I = 0; I < 2; ++I)
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        const int i0 = 8 * I + (I == 1 ? (pi & 3) : pi), c0 = 4 * kt + t;
        const int i = i0 >= c0 ? i0 : c0, c = i0 >= c0 ? c0 : i0;
        const int pnl = i >> 2;
        oRs[I][kt] = (pnl == 0 ? 0 : (pnl == 1 ? 16 : 48)) + 4 * c + (i & 3);
      }
    const int lc = lane < 24 ? lane : 0;
    const int j0 = lc < 12 ? 0 : 6;
    double acr[6];  // this lane's constraint row (6 nonzeros)
#pragma unroll
    for (int j = 0; j < 6; ++j) acr[j] = cAc[lc * 12 + j0 + j];
    double ng_ = 0.0, nb_ = 0.0, nd_ = 0.0, nm_ = 0.0, smu = 0.0;
    acc[0] = acc[1] = acc[2] = acc[3] = acc[4] = acc[5] = 0.0;
    if (k0 > N) return;
    tiles_begin<kT6>();
    if (k0 < N) prefetch_G<kT6>(k0, 0, 7);
    prefetch_R<kT6>(k0, 0);
    S6raw raw = load_s6(k0, do_update);
    double pp[3] = {0.0, 0.0, 0.0};  // updated pi_{k-1}, fragment form
    int b = 0;
#if SRBD_K3_UNROLL_RES
#pragma unroll 2
#endif
    for (int k = k0; k <= N; k += kstep, b ^= 1) {
      const int nu = k < N ? 12 : 0, nx = k > 0 ? 12 : 0, n = nu + nx;
      tiles_wait<kT6>(b);
      __syncwarp();
      set_bufs(b);
      const S6v cur = updated(raw, do_update, sp, sd);
      if (kTeam) {   // pi_{k-1}, updated like the warp of stage k - 1 does (same expression as in updated())
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          pp[j] = raw.ppi[j];
          if (do_update) pp[j] += sd * raw.pdpi[j];
        }
      }
      if (k + kstep <= N) {
        const int kn = k + kstep;
        tiles_begin<kT6>();
        if (kn < N) prefetch_G<kT6>(kn, b ^ 1, 7);
        prefetch_R<kT6>(kn, b ^ 1);
        raw = load_s6(kn, do_update);
        if (kn + 1 < N) prefetch_L2(kn + 1, false, false, true);
      }
      double* zb = sm + (b ? v2::wSX : v2::wSG);       // z (24)
      double* lb = sm + (b ? v2::wqx : v2::wQX);       // lam (24)
      double* pb = sm + v2::wXN + (b ? 24 : 0);        // pi (12) | x_{k+1} (12)   (wXN..wDI: 48 doubles)
      if (lane < 24) zb[lane] = lane < n ? cur.z : 0.0;
      if (k < N) {
        if (lane < 12) { pb[lane] = cur.pi; pb[12 + lane] = cur.xn; }
        if (lane < 24) lb[lane] = cur.lam;
      }
      if (do_update) {
        if (lane < n) zc(k, zs ^ 1)[0] = cur.z;
        if (k < N) {
          if (lane < 12) pic(k, zs ^ 1)[0] = cur.pi;
          if (lane < 24) { wsc(k, v2::oT)[0] = cur.t; wsc(k, v2::oLAM)[0] = cur.lam; }
        }
      }
      __syncwarp();
      // ---- fragment form: res_g = H z + g + G pi - [0; pi_{k-1}] - D^T lam ;  res_b = G^T z + b - x_{k+1} ----------
      double zk[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) zk[j] = zb[4 * j + t];
      double c0[2] = {sR[108 + t], sR[112 + t]}, c1[2] = {sR[116 + t], sR[120 + t]}, c2[2] = {sR[124 + t], sR[128 + t]};
      if (n < 24) { c1[1] = 0.0; c2[0] = 0.0; c2[1] = 0.0; }
      // dense lower block of rows 0..11 (R_k, or Q_N at stage N)
#pragma unroll
#if SRBD_K3_MERGE_HALVES
      // rows 8..11 of H z (first accumulator half) and rows 8..11 of G^T z (res_b, second half) multiply the same z:
      // one DMMA per k-tile for both (at stage N the second half is unused)
      const int pbr = ((k > 0 || kCG) ? 6 : 3) * v2::kGP + 4 * t;  // the b row of BAbt: row n (kCG: always row 24)
      double b0[2] = {sG[pbr], sG[pbr + 16]}, b1[2] = {sG[pbr + 32], 0.0};
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        dmma(c0[0], c0[1], zk[kt], sR[oRs[0][kt]], c0[0], c0[1]);
        const double r1 = sR[oRs[1][kt]], g1 = sG[oGT + 16 + v2::kGP * kt];
        dmma(c1[0], b1[0], zk[kt], pi < 4 ? r1 : g1, c1[0], b1[0]);
      }
#else
      for (int kt = 0; kt < 3; ++kt) {
        dmma(c0[0], c0[1], zk[kt], sR[oRs[0][kt]], c0[0], c0[1]);
        const double r1 = sR[oRs[1][kt]];
        dmma(c1[0], c1[1], zk[kt], pi < 4 ? r1 : 0.0, c1[0], c1[1]);
      }
#endif
      if (want_obj) {   // rows 0..11: c = g + H z so far; rows 12..23 (0 < k < N): H = diag(Q)
        double o = 0.5 * (zk[0] * (c0[0] + sR[108 + t]) + zk[1] * (c0[1] + sR[112 + t]) + zk[2] * (c1[0] + sR[116 + t]));
        if (k > 0 && k < N)
          o += zk[3] * fma(0.5 * cQ[t], zk[3], sR[120 + t]) + zk[4] * fma(0.5 * cQ[4 + t], zk[4], sR[124 + t]) +
               zk[5] * fma(0.5 * cQ[8 + t], zk[5], sR[128 + t]);
        if (r == 0) ob += o;
      }
      if (k < N) {
        if (k > 0) {  // diag(Q) on the x rows 12..23
          c1[1] = fma(cQ[t], zk[3], c1[1]);
          c2[0] = fma(cQ[4 + t], zk[4], c2[0]);
          c2[1] = fma(cQ[8 + t], zk[5], c2[1]);
        }
        double pk_[3], nl[6];
#pragma unroll
        for (int j = 0; j < 3; ++j) pk_[j] = pb[4 * j + t];
#pragma unroll
        for (int j = 0; j < 6; ++j) nl[j] = 0.0 - lb[4 * j + t];
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {  // + G pi
          dmma(c0[0], c0[1], pk_[kt], sG[gB + 16 * kt], c0[0], c0[1]);
          dmma(c1[0], c1[1], pk_[kt], sG[gB + 2 * v2::kGP + 16 * kt], c1[0], c1[1]);
          dmma(c2[0], c2[1], pk_[kt], sG[gB + 4 * v2::kGP + 16 * kt], c2[0], c2[1]);
        }
        if (k > 0) { c1[1] -= pp[0]; c2[0] -= pp[1]; c2[1] -= pp[2]; }
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {  // J^T (lam_u - lam_l) = -D^T lam
#if SRBD_K3_MERGE_HALVES
          // the rows of leg 1 (kt >= 3) only reach columns 6..11 = the second half of tile 0 and the first half of
          // tile 1: one DMMA whose first-half lanes supply Ac[.][8 + pi] and whose second-half lanes Ac[.][pi]
          if (kt < 3) dmma(c0[0], c0[1], nl[kt], cAc[oDt + 48 * kt], c0[0], c0[1]);
          else dmma(c1[0], c0[1], nl[kt], cAc[oDt + 48 * kt + (pi < 4 ? 8 : 0)], c1[0], c0[1]);
#else
          dmma(c0[0], c0[1], nl[kt], cAc[oDt + 48 * kt], c0[0], c0[1]);
          if (kt >= 3) dmma(c1[0], c1[1], nl[kt], cAc[oDt + 48 * kt + 8], c1[0], c1[1]);
#endif
        }
        // res_b
#if SRBD_K3_MERGE_HALVES
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
          dmma(b0[0], b0[1], zk[kt], sG[oGT + v2::kGP * kt], b0[0], b0[1]);  // stage 0: z[12..23] = 0
          if (kt >= 3) dmma(b1[0], b1[1], zk[kt], sG[oGT + 32 + v2::kGP * kt], b1[0], b1[1]);
        }
#else
        const int pbr = ((k > 0 || kCG) ? 6 : 3) * v2::kGP + 4 * t;  // the b row of BAbt: row n (kCG: always row 24)
        double b0[2] = {sG[pbr], sG[pbr + 16]}, b1[2] = {sG[pbr + 32], 0.0};
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
          dmma(b0[0], b0[1], zk[kt], sG[oGT + v2::kGP * kt], b0[0], b0[1]);  // stage 0: z[12..23] = 0
          dmma(b1[0], b1[1], zk[kt], sG[oGT + 32 + v2::kGP * kt], b1[0], b1[1]);
        }
#endif
        b0[0] -= pb[12 + t]; b0[1] -= pb[16 + t]; b1[0] -= pb[20 + t];
        if (r == 0) {
          wsf(k, v2::oRB)[0] = b0[0]; wsf(k, v2::oRB)[4] = b0[1]; wsf(k, v2::oRB)[8] = b1[0];
        }
        nb_ = amax_nan(amax_nan(amax_nan(nb_, b0[0]), b0[1]), b1[0]);
        pp[0] = pk_[0]; pp[1] = pk_[1]; pp[2] = pk_[2];
      } else {
        c0[0] -= pp[0]; c0[1] -= pp[1]; c1[0] -= pp[2];  // stage N: the x rows are rows 0..11
      }
      if (r == 0) {
        wsf(k, v2::oRG)[0] = c0[0]; wsf(k, v2::oRG)[4] = c0[1]; wsf(k, v2::oRG)[8] = c1[0];
        if (n == 24) {
          wsf(k, v2::oRG)[12] = c1[1]; wsf(k, v2::oRG)[16] = c2[0]; wsf(k, v2::oRG)[20] = c2[1];
        }
      }
      ng_ = amax_nan(amax_nan(amax_nan(ng_, c0[0]), c0[1]), c1[0]);
      if (n == 24) ng_ = amax_nan(amax_nan(amax_nan(ng_, c1[1]), c2[0]), c2[1]);
      // ---- constraint rows (row-per-lane) ------------------------------------------------------------------------------
      if (k < N) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) v += acr[j] * zb[j0 + j];
        const double rd = ((cur.lo - v) + cur.t) * cur.mk;
        const double rm = (cur.lam * cur.t) * cur.mk;
        if (lane < 24) {
          wsc(k, v2::oRD)[0] = rd;
          wsc(k, v2::oRM)[0] = rm;
          smu += rm;
          nd_ = amax_nan(nd_, rd);
          nm_ = amax_nan(nm_, rm);
        }
      }
    }
    acc[0] = ng_; acc[1] = nb_; acc[2] = nd_; acc[3] = nm_; acc[4] = smu; acc[5] = ob;
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------------
  // Exports of the hpipm-cpp facade (ocp_qp_ipm_solver.cpp:337-387), row-per-lane from the workspace of the LAST
  // factorization, same formulas as the generic kernel's write_outputs (ipm_solve.cuh).  Not on the throughput path.
  // ------------------------------------------------------------------------------------------------
  __device__ __forceinline__ void stat_row(int row, int col0, const double* v, int n) const {
    if (kExp && p.stat && row < p.stat_rows && lane == 0) {
      double* rr = p.stat + ((size_t)q * p.stat_rows + row) * SRBD_STAT_M;
      for (int i = 0; i < n; ++i) rr[col0 + i] = v[i];
    }
  }
  // factor panels pn (3 x kPanF, layout of oFT): T_pp = L_pp^-T in rows 4pp..4pp+3 of panel pp, L[i][j] (i >= 4 (j/4) + 4) in
  // row i of panel j / 4
  static __device__ __forceinline__ double pnT(const double* pn, int pp, int a, int b) { return pn[pp * v2::kPanF + (4 * pp + a) * 4 + b]; }
  static __device__ __forceinline__ double pnL(const double* pn, int i, int j) { return pn[(j >> 2) * v2::kPanF + i * 4 + (j & 3)]; }
  // v <- L^-T v (blocked back substitution)
  static __device__ void tri_bwd(const double* pn, double v[12]) {
#pragma unroll
    for (int pp = 2; pp >= 0; --pp) {
      double b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double a = v[4 * pp + j];
        for (int i = 4 * pp + 4; i < 12; ++i) a -= pnL(pn, i, 4 * pp + j) * v[i];
        b[j] = a;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        double y = 0.0;
#pragma unroll
        for (int j = a; j < 4; ++j) y += pnT(pn, pp, a, j) * b[j];
        v[4 * pp + a] = y;
      }
    }
  }
  // v <- L^-1 v (blocked forward substitution; L_pp^-1 = T_pp^T)
  static __device__ void tri_fwd(const double* pn, double v[12]) {
#pragma unroll
    for (int pp = 0; pp < 3; ++pp) {
      double b[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        double acc = v[4 * pp + a];
        for (int j = 0; j < 4 * pp; ++j) acc -= pnL(pn, 4 * pp + a, j) * v[j];
        b[a] = acc;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        double y = 0.0;
#pragma unroll
        for (int j = 0; j <= a; ++j) y += pnT(pn, pp, j, a) * b[j];
        v[4 * pp + a] = y;
      }
    }
  }
  __device__ __noinline__ void export_riccati(bool unc) {
    double* sPn = sm;          // factor panels of the stage (306)
    double* sPk = sm + 320;    // P_k (144)
    double* sx_ = sm + 480;    // z_k (24)
    double* sK = sm + 512;     // K_k, column-major (144)
    const size_t qS = (size_t)q * (N + 1), qN = (size_t)q * N;
    const int l0 = lane < 24 ? lane : 0;
    for (int k = 1; k <= N; ++k) {
      const int nu = k < N ? 12 : 0;
      __syncwarp();
      for (int e = lane; e < 144; e += 32) {
        const double v = ws_ld(ws(k - 1, v2::oP) + e);
        sPk[e] = v;
        p.ric_P[(qS + k) * 144 + e] = v;   // symmetric: row-major == column-major
      }
      if (k < N)
        for (int e = lane; e < 306; e += 32) sPn[e] = ws_ld(ws(k, v2::oFT) + e);
      if (lane < 24) sx_[lane] = lane < nu + 12 ? ws_ld(zc(k, zs)) : 0.0;
      __syncwarp();
      if (lane < 12) {
        double v;
        if (unc) v = ws_ld(ws(k, v2::oPV) + lane);
        else {   // p_k = pi_k - P_k x_k
          double acc = 0.0;
          for (int j = 0; j < 12; ++j) acc += sPk[lane * 12 + j] * sx_[nu + j];
          v = ws_ld(pic(k - 1, zs) - l0 + lane) - acc;
        }
        p.ric_p[(qS + k) * 12 + lane] = v;
      }
      if (k < N) {
        if (lane < 13) {   // lanes 0..11: column c = lane of K_k = -L^-T Ls^T; lane 12: -L^-T lv
          double y[12];
#pragma unroll
          for (int j = 0; j < 12; ++j) y[j] = sPn[(j >> 2) * v2::kPanF + (lane < 12 ? 12 + lane : 24) * 4 + (j & 3)];
          tri_bwd(sPn, y);
          if (lane < 12) {
#pragma unroll
            for (int i = 0; i < 12; ++i) {
              sK[i + 12 * lane] = -y[i];
              p.ric_K[(qN + k) * 144 + i + 12 * lane] = -y[i];
            }
          } else if (unc) {
#pragma unroll
            for (int i = 0; i < 12; ++i) p.ric_k[(qN + k) * 12 + i] = -y[i];
          }
        }
        __syncwarp();
        if (!unc && lane < 12) {   // k_k = u_k - K_k x_k
          double acc = 0.0;
          for (int c = 0; c < 12; ++c) acc += sK[lane + 12 * c] * sx_[12 + c];
          p.ric_k[(qN + k) * 12 + lane] = sx_[lane] - acc;
        }
      }
    }
    __syncwarp();
    // ---- stage 0 (ocp_qp_ipm_solver.cpp:349-373) from the raw stage-0 blocks --------------------------------------------
    {
      const double* raw = p.raw0 + (size_t)q * p.raw0_stride;
      const double *A0 = raw, *B0 = raw + 144, *b0 = raw + 288, *S0 = raw + 300, *Q0 = raw + 444, *q0 = raw + 588;
      const double* x0 = p.x0 + (size_t)q * 12;
      double* sP1 = sPk;            // P_1
      double* H0 = sm + 672;        // [i * 12 + j], nu x nx
      double* AtP = sm + 816;
      double* GH = sm + 960;
      double* BtP = sm + 1104;
      double* sv = sm + 1248;       // k0 (12) | p1 (12)
      for (int e = lane; e < 144; e += 32) sP1[e] = ws_ld(ws(0, v2::oP) + e);
      for (int e = lane; e < 306; e += 32) sPn[e] = ws_ld(ws(0, v2::oFT) + e);
      if (lane < 12) sx_[lane] = ws_ld(zc(0, zs));   // u_0
      __syncwarp();
      if (p.ric_Lr0 && lane < 12) {
        // column c = lane of Lr_0 from the factor panels: the rows below the diagonal block are stored as they are, the
        // diagonal block L_pp is the inverse of X = T_pp^T (T_pp = L_pp^-T is what the blocked solves use)
        const int pp = lane >> 2, c = lane & 3;
        double* o = p.ric_Lr0 + (size_t)q * 144 + 12 * lane;
        for (int i = 0; i < 4 * pp; ++i) o[i] = 0.0;
        double Lc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int a = c; a < 4; ++a) {
          double acc = a == c ? 1.0 : 0.0;
          for (int m = c; m < a; ++m) acc -= pnT(sPn, pp, m, a) * Lc[m];
          const double d = pnT(sPn, pp, a, a);
          Lc[a] = d != 0.0 ? acc / d : 0.0;   // (a failed pivot left a zero in the block inverse)
        }
        for (int a = 0; a < 4; ++a) o[4 * pp + a] = a >= c ? Lc[a] : 0.0;
        for (int i = 4 * pp + 4; i < 12; ++i) o[i] = pnL(sPn, i, lane);
      }
      for (int e = lane; e < 144; e += 32) {
        const int i = e / 12, j = e % 12;
        double a1 = 0.0, a2 = 0.0;
        for (int l = 0; l < 12; ++l) {
          a1 += B0[l + 12 * i] * sP1[l * 12 + j];
          a2 += A0[l + 12 * i] * sP1[l * 12 + j];
        }
        BtP[e] = a1; AtP[e] = a2;
      }
      __syncwarp();
      for (int e = lane; e < 144; e += 32) {
        const int i = e / 12, j = e % 12;
        double acc = 0.0;
        for (int l = 0; l < 12; ++l) acc += BtP[i * 12 + l] * A0[l + 12 * j];
        H0[e] = S0[i + 12 * j] + acc;
      }
      __syncwarp();
      if (lane < 12) {   // column c = lane of GH = L^-T L^-1 H0
        double y[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) y[i] = H0[i * 12 + lane];
        tri_fwd(sPn, y);
        tri_bwd(sPn, y);
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          GH[i * 12 + lane] = y[i];
          p.ric_K[qN * 144 + i + 12 * lane] = -y[i];
        }
      }
      __syncwarp();
      if (lane < 12) {
        double acc = 0.0;
        for (int j = 0; j < 12; ++j) acc += -GH[lane * 12 + j] * x0[j];
        const double k0 = sx_[lane] - acc;
        sv[lane] = k0;
        p.ric_k[qN * 12 + lane] = k0;
        sv[12 + lane] = p.ric_p[(qS + 1) * 12 + lane];   // (written by this lane above)
      }
      __syncwarp();
      double* P0 = p.ric_P + qS * 144;
      for (int e = lane; e < 144; e += 32) {
        const int i = e % 12, j = e / 12;  // column-major output
        double s1 = 0.0, s2 = 0.0;
        for (int l = 0; l < 12; ++l) s1 += H0[l * 12 + i] * GH[l * 12 + j];
        for (int l = 0; l < 12; ++l) s2 += AtP[i * 12 + l] * A0[l + 12 * j];
        const double v = (Q0[i + 12 * j] - s1) + s2;
        P0[e] = v;
        sK[e] = v;
      }
      __syncwarp();
      if (lane < 12) {
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int l = 0; l < 12; ++l) s1 += A0[l + 12 * lane] * sv[12 + l];
        for (int l = 0; l < 12; ++l) s2 += AtP[lane * 12 + l] * b0[l];
        for (int l = 0; l < 12; ++l) s3 += H0[l * 12 + lane] * sv[l];
        const double p0 = ((q0[lane] + s1) + s2) + s3;
        p.ric_p[qS * 12 + lane] = p0;
        double acc = 0.0;
        for (int j = 0; j < 12; ++j) acc += sK[lane + 12 * j] * x0[j];
        p.sol_pi[qS * 12 + lane] = p0 + acc;
      }
      __syncwarp();
    }
  }


  // ------------------------------------------------------------------------------------------------
  // S6 + S1 FUSED (throughput instantiations): the residual sweep runs BACKWARD inside the factorization sweep.  The
  // residuals of a stage (res_g, res_b, res_d, res_m) only need the iterate of the stage and of its two neighbours, and the
  // factorization of stage k consumes exactly those residuals: one pass over the BAbt record, the R tile and the stage
  // vectors instead of two (15 % less DRAM traffic per iteration, no second tile pipeline).  The convergence test needs the
  // norms of ALL stages, i.e. the end of the sweep: the factorization of the last (converged) iterate is wasted, one in
  // thirteen.  Element by element the same operations as residual_sweep + sweep_factor; the duality measure is re-summed
  // from the stored res_m in forward stage order, so the results are bit-identical to the two separate sweeps.
  // (Not for the exporting instantiation: hpipm's getters return the factorization the last STEP was computed with.)
  // ------------------------------------------------------------------------------------------------
  struct RFraw { double z, pi, lam, t, lo, mk, dz, dpi, dlam, dt, ppi[3], pdpi[3]; };
  __device__ __forceinline__ RFraw load_rf(int k, bool do_update) const {
    RFraw v;
    v.z = ws_ld(wsc(k, v2::oZ));
    v.pi = 0.0; v.lam = 0.0; v.t = 1.0; v.lo = 0.0; v.mk = 0.0;
    v.dz = 0.0; v.dpi = 0.0; v.dlam = 0.0; v.dt = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) { v.ppi[j] = 0.0; v.pdpi[j] = 0.0; }
    if (k > 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) v.ppi[j] = ws_ld(wsf(k - 1, v2::oPI) + 4 * j);
    }
    if (k < N) {
      v.pi = ws_ld(wsc(k, v2::oPI)); v.lam = ws_ld(wsc(k, v2::oLAM)); v.t = ws_ld(wsc(k, v2::oT));
      v.lo = __ldg(gDL(k)); v.mk = mask_of(k);
    }
    if (do_update) {
      v.dz = ws_ld(wsc(k, v2::oDZ));
      if (k > 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) v.pdpi[j] = ws_ld(wsf(k - 1, v2::oDPI) + 4 * j);
      }
      if (k < N) {
        v.dpi = ws_ld(wsc(k, v2::oDPI));
        v.dt = ws_ld(wsc(k, v2::oDT));
        v.dlam = ws_ld(wsc(k, v2::oDLAM));
      }
    }
    return v;
  }
  __device__ void sweep_resfac(double res[4], double& mu, int nc_mask, bool do_update, double sp, double sd) {
    constexpr bool kT1 = (kTma & 1) != 0;
    const double reg = p.a.reg_prim;
    double* sP = sm + v2::wP;
    double* sPan = sm + v2::wPAN;
    const int r = fr, t = ft, pi = fpi;
    const int gA = (r >> 2) * v2::kGP + 4 * t + (r & 3);
    const int gB = (pi >> 2) * v2::kGP + 4 * t + (pi & 3);
    const int prA0 = (25 + r) * 4 + t, prA1 = (r < 4 ? 33 + r : 8 + r) * 4 + t, prA2 = (16 + r) * 4 + t;
    const int prB0 = (25 + pi) * 4 + t, prB1 = (pi < 4 ? 33 + pi : 8 + pi) * 4 + t, prB2 = (16 + pi) * 4 + t;
    const int prE0 = r * 4 + t;
    const int oDt = t * 12 + pi;  // Ac[4kt+t][8I+pi] : + 8 I + 48 kt
    const int oGT = 4 * pi + t;   // G[4kt+t][8I+pi]  : + 32 I + kGP kt
    const bool dg0 = (r == t), dg1 = (r == 4 + t);
    const double regd0 = dg0 ? reg : 0.0, regd1 = dg1 ? reg : 0.0;
    const int rel0 = ((r >> 2) ? 16 : 0) + (r & 3) + 4 * t;
    const int rel1 = 48 + (r & 3) + 4 * t;
    const double qd3 = dg1 ? cQ[t] : 0.0, qd4 = dg0 ? cQ[4 + t] : 0.0, qd5 = dg1 ? cQ[8 + t] : 0.0;
    const int lc = lane < 24 ? lane : 0;
    const int j0 = lc < 12 ? 0 : 6;
    double ng_ = 0.0, nb_ = 0.0, nd_ = 0.0, nm_ = 0.0;
    // symmetric R block (Q_N at stage N) as B fragments from the lower-triangle tile: element (8I+pi, 4kt+t)
    auto oRs = [&](int I, int kt) {
      const int i0 = 8 * I + (I == 1 ? (pi & 3) : pi), c0_ = 4 * kt + t;
      const int i = i0 >= c0_ ? i0 : c0_, c = i0 >= c0_ ? c0_ : i0;
      const int pnl = i >> 2;
      return (pnl == 0 ? 0 : (pnl == 1 ? 16 : 48)) + 4 * c + (i & 3);
    };
    // ---- prologue: the R tile of stage N into buffer 1, stage N-1 into buffer 0 ------------------------------------------
    tiles_begin<kT1>();
    prefetch_R<kT1>(N, 1);
    prefetch_G<kT1>(N - 1, 0, 7);
    prefetch_R<kT1>(N - 1, 0);
    RFraw raw = load_rf(N, do_update);
    tiles_wait<kT1>(1);
    if (kT1) tiles_wait<kT1>(0);   // (cp.async: one wait covers both)
    __syncwarp();
    double pk[3];
    {
      // stage N: res_g = Q_N z + q_N - pi_{N-1};  P_N = Q_N + reg I, p_N = res_g
      set_bufs(1);
      double zN = raw.z;
      if (do_update) zN += sp * raw.dz;
      double pp[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) { pp[j] = raw.ppi[j]; if (do_update) pp[j] += sd * raw.pdpi[j]; }
      raw = load_rf(N - 1, do_update);
      double* zb = sm + v2::wSX;
      double* pb0 = sm + v2::wXN;   // the vector buffer of trip N-1 (b = 0): [pi | x_{k+1}]
      if (lane < 24) zb[lane] = lane < 12 ? zN : 0.0;
      if (lane < 12) pb0[12 + lane] = zN;
      if (do_update && lane < 12) wsc(N, v2::oZ)[0] = zN;
      __syncwarp();
      double zk[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) zk[j] = zb[4 * j + t];
      double c0[2] = {sR[108 + t], sR[112 + t]}, c1[2] = {sR[116 + t], sR[120 + t]};
      c1[1] = 0.0;
#if SRBD_K3_MERGE_HALVES
      double bdum = 0.0;
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        dmma(c0[0], c0[1], zk[kt], sR[oRs(0, kt)], c0[0], c0[1]);
        const double r1 = sR[oRs(1, kt)], g1 = sG[oGT + 16 + v2::kGP * kt];
        dmma(c1[0], bdum, zk[kt], pi < 4 ? r1 : g1, c1[0], bdum);
      }
#else
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        dmma(c0[0], c0[1], zk[kt], sR[oRs(0, kt)], c0[0], c0[1]);
        const double r1 = sR[oRs(1, kt)];
        dmma(c1[0], c1[1], zk[kt], pi < 4 ? r1 : 0.0, c1[0], c1[1]);
      }
#endif
      c0[0] -= pp[0]; c0[1] -= pp[1]; c1[0] -= pp[2];
      if (r == 0) { wsf(N, v2::oRG)[0] = c0[0]; wsf(N, v2::oRG)[4] = c0[1]; wsf(N, v2::oRG)[8] = c1[0]; }
      ng_ = amax_nan(amax_nan(amax_nan(ng_, c0[0]), c0[1]), c1[0]);
      pk[0] = c0[0]; pk[1] = c0[1]; pk[2] = c1[0];
      const double* rs = gRecL(N) - kLaneOff * lane;  // Q_N: the R tile of the last stage record
      if (lane < 12) {
#pragma unroll
        for (int c = 0; c < 12; ++c) {
          if (c <= lane) {
            double v = __ldg(rs + ((lane >> 2) == 0 ? 0 : ((lane >> 2) == 1 ? 16 : 48)) + 4 * c + (lane & 3));
            if (c == lane) v += reg;
            sP[lane * 12 + c] = v;
            sP[c * 12 + lane] = v;
          }
        }
      }
      __syncwarp();
      for (int e = lane; e < 144; e += 32) ws(N - 1, v2::oP)[e] = sP[e];
      if (r == 0) {
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) wsf(N, v2::oPV)[4 * kt] = pk[kt];
      }
    }
    for (int k = N - 1; k >= 0; --k) {
      const int b = (N - 1 - k) & 1;
      const bool xr = k > 0;
      const int n = xr ? 24 : 12;
      if (k < N - 1) { tiles_wait<kT1>(b); }
      __syncwarp();
      set_bufs(b);
      // ---- the iterate of stage k (d_update_var_qp on the prefetched vectors) ------------------------------------------
      double z = raw.z, piv = raw.pi, lam = raw.lam, tt = raw.t;
      const double lo = raw.lo, mk = raw.mk;
      double pp[3] = {raw.ppi[0], raw.ppi[1], raw.ppi[2]};
      if (do_update) {
        z += sp * raw.dz;
        piv += sd * raw.dpi;
        tt += sp * raw.dt;
        lam += sd * raw.dlam;
        if (p.a.t_lam_min == 2 && mk != 0.0) {
          tt = tt < p.a.t_min ? p.a.t_min : tt;
          lam = lam < p.a.lam_min ? p.a.lam_min : lam;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) pp[j] += sd * raw.pdpi[j];
      }
      if (k > 0) {
        tiles_begin<kT1>();
        prefetch_G<kT1>(k - 1, b ^ 1, 7);
        prefetch_R<kT1>(k - 1, b ^ 1);
        raw = load_rf(k - 1, do_update);
      }
      double* zb = sm + (b ? v2::wSX : v2::wSG);       // z (24)
      double* lb = sm + (b ? v2::wQX : v2::wqx);       // lam (24)
      double* pb = sm + v2::wXN + (b ? 24 : 0);        // pi (12) | x_{k+1} (12, written by the previous trip)
      double* pbn = sm + v2::wXN + (b ? 0 : 24);
      if (lane < 24) { zb[lane] = lane < n ? z : 0.0; lb[lane] = lam; }
      if (lane < 12) pb[lane] = piv;
      if (xr && lane >= 12 && lane < 24) pbn[lane] = z;   // x_k for trip k - 1
      if (do_update) {
        if (lane < n) wsc(k, v2::oZ)[0] = z;
        if (lane < 12) wsc(k, v2::oPI)[0] = piv;
        if (lane < 24) { wsc(k, v2::oT)[0] = tt; wsc(k, v2::oLAM)[0] = lam; }
      }
      __syncwarp();
      // ---- residuals, fragment form: res_g = H z + g + G pi - [0; pi_{k-1}] - D^T lam ;  res_b = G^T z + b - x_{k+1} -------
      double c0[2], c1[2], c2[2], rbv[3], rd, rm;
      {
        double zk[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) zk[j] = zb[4 * j + t];
        c0[0] = sR[108 + t]; c0[1] = sR[112 + t]; c1[0] = sR[116 + t]; c1[1] = sR[120 + t]; c2[0] = sR[124 + t]; c2[1] = sR[128 + t];
        if (!xr) { c1[1] = 0.0; c2[0] = 0.0; c2[1] = 0.0; }
        const int pbr = ((xr || kCG) ? 6 : 3) * v2::kGP + 4 * t;  // the b row of BAbt: row n (kCG: always row 24)
        double b0[2] = {sG[pbr], sG[pbr + 16]}, b1[2] = {sG[pbr + 32], 0.0};
#if SRBD_K3_MERGE_HALVES
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(c0[0], c0[1], zk[kt], sR[oRs(0, kt)], c0[0], c0[1]);
          const double r1 = sR[oRs(1, kt)], g1 = sG[oGT + 16 + v2::kGP * kt];
          dmma(c1[0], b1[0], zk[kt], pi < 4 ? r1 : g1, c1[0], b1[0]);
        }
#else
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(c0[0], c0[1], zk[kt], sR[oRs(0, kt)], c0[0], c0[1]);
          const double r1 = sR[oRs(1, kt)];
          dmma(c1[0], c1[1], zk[kt], pi < 4 ? r1 : 0.0, c1[0], c1[1]);
        }
#endif
        if (xr) {  // diag(Q) on the x rows 12..23
          c1[1] = fma(cQ[t], zk[3], c1[1]);
          c2[0] = fma(cQ[4 + t], zk[4], c2[0]);
          c2[1] = fma(cQ[8 + t], zk[5], c2[1]);
        }
        double pk_[3], nl[6];
#pragma unroll
        for (int j = 0; j < 3; ++j) pk_[j] = pb[4 * j + t];
#pragma unroll
        for (int j = 0; j < 6; ++j) nl[j] = 0.0 - lb[4 * j + t];
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {  // + G pi
          dmma(c0[0], c0[1], pk_[kt], sG[gB + 16 * kt], c0[0], c0[1]);
          dmma(c1[0], c1[1], pk_[kt], sG[gB + 2 * v2::kGP + 16 * kt], c1[0], c1[1]);
          dmma(c2[0], c2[1], pk_[kt], sG[gB + 4 * v2::kGP + 16 * kt], c2[0], c2[1]);
        }
        if (xr) { c1[1] -= pp[0]; c2[0] -= pp[1]; c2[1] -= pp[2]; }
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {  // J^T (lam_u - lam_l) = -D^T lam
#if SRBD_K3_MERGE_HALVES
          if (kt < 3) dmma(c0[0], c0[1], nl[kt], cAc[oDt + 48 * kt], c0[0], c0[1]);
          else dmma(c1[0], c0[1], nl[kt], cAc[oDt + 48 * kt + (pi < 4 ? 8 : 0)], c1[0], c0[1]);
#else
          dmma(c0[0], c0[1], nl[kt], cAc[oDt + 48 * kt], c0[0], c0[1]);
          if (kt >= 3) dmma(c1[0], c1[1], nl[kt], cAc[oDt + 48 * kt + 8], c1[0], c1[1]);
#endif
        }
#if SRBD_K3_MERGE_HALVES
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
          dmma(b0[0], b0[1], zk[kt], sG[oGT + v2::kGP * kt], b0[0], b0[1]);  // stage 0: z[12..23] = 0
          if (kt >= 3) dmma(b1[0], b1[1], zk[kt], sG[oGT + 32 + v2::kGP * kt], b1[0], b1[1]);
        }
#else
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
          dmma(b0[0], b0[1], zk[kt], sG[oGT + v2::kGP * kt], b0[0], b0[1]);
          dmma(b1[0], b1[1], zk[kt], sG[oGT + 32 + v2::kGP * kt], b1[0], b1[1]);
        }
#endif
        b0[0] -= pb[12 + t]; b0[1] -= pb[16 + t]; b1[0] -= pb[20 + t];
        rbv[0] = b0[0]; rbv[1] = b0[1]; rbv[2] = b1[0];
        if (r == 0) {
          wsf(k, v2::oRB)[0] = b0[0]; wsf(k, v2::oRB)[4] = b0[1]; wsf(k, v2::oRB)[8] = b1[0];
          wsf(k, v2::oRG)[0] = c0[0]; wsf(k, v2::oRG)[4] = c0[1]; wsf(k, v2::oRG)[8] = c1[0];
          if (xr) { wsf(k, v2::oRG)[12] = c1[1]; wsf(k, v2::oRG)[16] = c2[0]; wsf(k, v2::oRG)[20] = c2[1]; }
        }
        nb_ = amax_nan(amax_nan(amax_nan(nb_, b0[0]), b0[1]), b1[0]);
        ng_ = amax_nan(amax_nan(amax_nan(ng_, c0[0]), c0[1]), c1[0]);
        if (xr) ng_ = amax_nan(amax_nan(amax_nan(ng_, c1[1]), c2[0]), c2[1]);
        // constraint rows (row-per-lane)
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) v += cAc[lc * 12 + j0 + j] * zb[j0 + j];
        rd = ((lo - v) + tt) * mk;
        rm = (lam * tt) * mk;
        if (lane < 24) {
          wsc(k, v2::oRD)[0] = rd;
          wsc(k, v2::oRM)[0] = rm;
          nd_ = amax_nan(nd_, rd);
          nm_ = amax_nan(nm_, rm);
        }
      }
      // ---- factorization of stage k (sweep_factor's body on the residuals just formed) -------------------------------------
      double* gbuf = sm + (b ? v2::wqx : v2::wQX);
      double* Gbuf = sm + (b ? v2::wSG : v2::wSX);
      if (lane < 24) {
        const double ti = 1.0 / tt;
        Gbuf[lane] = (ti * lam) * mk;
        gbuf[lane] = (ti * (rm - lam * rd)) * mk;
      }
      __syncwarp();
      double GF[3][3], GPF[3][3], PPF[2][3];
      {
        const bool okA1 = xr || r < 4, okB1 = xr || pi < 4;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          GF[0][kt] = sG[gA + 16 * kt];
          const double a1 = sG[gA + 2 * v2::kGP + 16 * kt], a2 = sG[gA + 4 * v2::kGP + 16 * kt];
          GF[1][kt] = okA1 ? a1 : 0.0;
          GF[2][kt] = xr ? a2 : 0.0;
          GPF[0][kt] = sG[gB + 16 * kt];
          const double b1_ = sG[gB + 2 * v2::kGP + 16 * kt], b2_ = sG[gB + 4 * v2::kGP + 16 * kt];
          GPF[1][kt] = okB1 ? b1_ : 0.0;
          GPF[2][kt] = xr ? b2_ : 0.0;
          PPF[0][kt] = sP[pi * 12 + 4 * kt + t];
          const double p1 = sP[(8 + (pi & 3)) * 12 + 4 * kt + t];
          PPF[1][kt] = pi < 4 ? p1 : 0.0;
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int e2 = lane + 32 * rr;
        if (e2 < 42) {
          const int leg = e2 >= 21 ? 1 : 0, e = e2 - 21 * leg;
          double acc = 0.0;
#pragma unroll
          for (int g = 0; g < 12; ++g) acc += Gbuf[12 * leg + g] * cW[(12 * leg + g) * v2::kW2 + e];
          sm[v2::wS + e2] = acc;
        }
      }
      {
        double d0[2] = {0.0, 0.0}, d1[2] = {0.0, 0.0};
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(d0[0], d0[1], rbv[kt], PPF[0][kt], d0[0], d0[1]);
          dmma(d1[0], d1[1], rbv[kt], PPF[1][kt], d1[0], d1[1]);
        }
        if (r == 0) {
          wsf(k, v2::oPRB)[0] = d0[0]; wsf(k, v2::oPRB)[4] = d0[1]; wsf(k, v2::oPRB)[8] = d1[0];
        }
        const double tk[3] = {d0[0] + pk[0], d0[1] + pk[1], d1[0] + pk[2]};
#pragma unroll
        for (int kt = 0; kt < 6; ++kt) {
          const double gam = gbuf[4 * kt + t];
#if SRBD_K3_MERGE_HALVES
          if (kt < 3) dmma(c0[0], c0[1], gam, cAc[oDt + 48 * kt], c0[0], c0[1]);
          else dmma(c1[0], c0[1], gam, cAc[oDt + 48 * kt + (pi < 4 ? 8 : 0)], c1[0], c0[1]);
#else
          dmma(c0[0], c0[1], gam, cAc[oDt + 48 * kt], c0[0], c0[1]);
          if (kt >= 3) dmma(c1[0], c1[1], gam, cAc[oDt + 48 * kt + 8], c1[0], c1[1]);
#endif
        }
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(c0[0], c0[1], tk[kt], GPF[0][kt], c0[0], c0[1]);
          dmma(c1[0], c1[1], tk[kt], GPF[1][kt], c1[0], c1[1]);
          dmma(c2[0], c2[1], tk[kt], GPF[2][kt], c2[0], c2[1]);
        }
      }
      double ALF[3][3];
#pragma unroll
      for (int I = 0; I < 3; ++I) {
        double a0 = 0.0, a1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          dmma(a0, a1, GF[I][kt], PPF[0][kt], a0, a1);
          dmma(e0, e1, GF[I][kt], PPF[1][kt], e0, e1);
        }
        ALF[I][0] = a0; ALF[I][1] = a1; ALF[I][2] = e0;
      }
      double MF[3][6];
#pragma unroll
      for (int I = 0; I < 3; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J) {
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int kt = 0; kt < 3; ++kt) dmma(a0, a1, ALF[I][kt], GPF[J][kt], a0, a1);
          MF[I][2 * J] = a0; MF[I][2 * J + 1] = a1;
        }
      __syncwarp();  // wS visible
      {
        double hf[5];
#pragma unroll
        for (int f = 0; f < 5; ++f) {
          const int I = f < 2 ? 0 : 1, pq = f < 2 ? f : f - 2;
          const bool valid = (8 * I + r < 12) && (4 * pq + t <= 8 * I + r);
          double v = 0.0;
          if (valid) v = sR[(I == 0 ? rel0 : rel1) + 16 * pq];
          if (offS[f] >= 0) v += sm[v2::wS + offS[f]];
          hf[f] = v;
        }
        MF[0][0] = (hf[0] + MF[0][0]) + regd0;
        MF[0][1] = (hf[1] + MF[0][1]) + regd1;
        MF[1][0] = hf[2] + MF[1][0];
        MF[1][1] = hf[3] + MF[1][1];
        MF[1][2] = (hf[4] + MF[1][2]) + regd0;
        MF[1][3] = ((xr ? qd3 : 0.0) + MF[1][3]) + regd1;
        MF[2][4] = ((xr ? qd4 : 0.0) + MF[2][4]) + regd0;
        MF[2][5] = ((xr ? qd5 : 0.0) + MF[2][5]) + regd1;
      }
#pragma unroll
      for (int pq = 0; pq < 3; ++pq) {
        double* pan = sPan + pq * v2::kPan;
        if (pq < 2) pan[prA0] = MF[0][pq];
        pan[prA1] = MF[1][pq];
        pan[prA2] = MF[2][pq];
        pan[prE0] = pq == 0 ? (dg0 ? 1.0 : 0.0) : ((pq == 1 && dg1) ? 1.0 : 0.0);
        if (r < 4) pan[(8 + r) * 4 + t] = (pq == 2 && dg0) ? 1.0 : 0.0;
        if (r == 0) pan[96 + t] = pq == 0 ? c0[0] : (pq == 1 ? c0[1] : c1[0]);
        __syncwarp();
        const double* dgb = pan + (25 + 4 * pq) * 4;
        const double a00 = dgb[0];
        const double2 a1x = *reinterpret_cast<const double2*>(dgb + 4);
        const double2 a2x = *reinterpret_cast<const double2*>(dgb + 8);
        const double a22 = dgb[10];
        const double2 a3x = *reinterpret_cast<const double2*>(dgb + 12);
        const double2 a3y = *reinterpret_cast<const double2*>(dgb + 14);
        const int prow = lane < 4 * pq ? 4 + lane : (lane < 12 ? 25 + lane : (lane < 25 ? lane : lane - 25));
        double2* own = reinterpret_cast<double2*>(pan + prow * 4);
        const double2 x01 = own[0], x23 = own[1];
        const double i0 = rsqrt_pivot(a00);
        const double L10 = a1x.x * i0, L20 = a2x.x * i0, L30 = a3x.x * i0;
        const double i1 = rsqrt_pivot(fma(-L10, L10, a1x.y));
        const double L21 = fma(-L20, L10, a2x.y) * i1, L31 = fma(-L30, L10, a3x.y) * i1;
        const double i2 = rsqrt_pivot(fma(-L21, L21, fma(-L20, L20, a22)));
        const double L32 = fma(-L31, L21, fma(-L30, L20, a3y.x)) * i2;
        const double i3 = rsqrt_pivot(fma(-L32, L32, fma(-L31, L31, fma(-L30, L30, a3y.y))));
        const double l0 = x01.x * i0;
        const double l1 = fma(-l0, L10, x01.y) * i1;
        const double l2 = fma(-l1, L21, fma(-l0, L20, x23.x)) * i2;
        const double l3 = fma(-l2, L32, fma(-l1, L31, fma(-l0, L30, x23.y))) * i3;
        own[0] = make_double2(l0, l1);
        own[1] = make_double2(l2, l3);
        __syncwarp();
        const double nlv = -pan[96 + t];
        double dum;
        if (pq == 0) {
          const double nA0 = -pan[prA0], nA1 = -pan[prA1], nA2 = -pan[prA2];
          const double B0 = pan[prB0], B1 = pan[prB1], B2 = pan[prB2];
          dmma(dum, MF[0][1], nA0, B0, MF[0][0], MF[0][1]);
          dmma(dum, MF[1][1], nA1, B0, MF[1][0], MF[1][1]);
          dmma(dum, MF[2][1], nA2, B0, MF[2][0], MF[2][1]);
          dmma(dum, c0[1], nlv, B0, c0[0], c0[1]);
          dmma(MF[1][2], MF[1][3], nA1, B1, MF[1][2], MF[1][3]);
          dmma(MF[2][2], MF[2][3], nA2, B1, MF[2][2], MF[2][3]);
          dmma(c1[0], c1[1], nlv, B1, c1[0], c1[1]);
          dmma(MF[2][4], MF[2][5], nA2, B2, MF[2][4], MF[2][5]);
          dmma(c2[0], c2[1], nlv, B2, c2[0], c2[1]);
        } else if (pq == 1) {
          const double nA1 = -pan[prA1], nA2 = -pan[prA2];
          const double B1 = pan[prB1], B2 = pan[prB2];
          dmma(MF[1][2], MF[1][3], nA1, B1, MF[1][2], MF[1][3]);
          dmma(MF[2][2], MF[2][3], nA2, B1, MF[2][2], MF[2][3]);
          dmma(c1[0], c1[1], nlv, B1, c1[0], c1[1]);
          dmma(MF[2][4], MF[2][5], nA2, B2, MF[2][4], MF[2][5]);
          dmma(c2[0], c2[1], nlv, B2, c2[0], c2[1]);
        } else {
          const double nA1 = -pan[prA1], nA2 = -pan[prA2];
          const double B1 = pan[prB1], B2 = pan[prB2];
          dmma(dum, MF[1][3], nA1, B1, MF[1][2], MF[1][3]);
          dmma(dum, MF[2][3], nA2, B1, MF[2][2], MF[2][3]);
          dmma(MF[2][4], MF[2][5], nA2, B2, MF[2][4], MF[2][5]);
          dmma(dum, c1[1], nlv, B1, c1[0], c1[1]);
          dmma(c2[0], c2[1], nlv, B2, c2[0], c2[1]);
        }
      }
      if (xr) {
        if (r >= 4 && t <= r - 4) { sP[(r - 4) * 12 + t] = MF[1][3]; sP[t * 12 + (r - 4)] = MF[1][3]; }
        sP[(4 + r) * 12 + t] = MF[2][3]; sP[t * 12 + (4 + r)] = MF[2][3];
        if (t <= r) { sP[(4 + r) * 12 + 4 + t] = MF[2][4]; sP[(4 + t) * 12 + 4 + r] = MF[2][4]; }
        if (4 + t <= r) { sP[(4 + r) * 12 + 8 + t] = MF[2][5]; sP[(8 + t) * 12 + 4 + r] = MF[2][5]; }
        pk[0] = c1[1]; pk[1] = c2[0]; pk[2] = c2[1];
        if (r == 0) {
          wsf(k, v2::oPV)[0] = pk[0]; wsf(k, v2::oPV)[4] = pk[1]; wsf(k, v2::oPV)[8] = pk[2];
        }
        __syncwarp();
        for (int e = lane; e < 72; e += 32)
          reinterpret_cast<double2*>(ws(k - 1, v2::oP))[e] = reinterpret_cast<const double2*>(sP)[e];
      }
#pragma unroll
      for (int pq = 0; pq < 3; ++pq) {
        const double2* src = reinterpret_cast<const double2*>(sPan + pq * v2::kPan);
        double2* dst = reinterpret_cast<double2*>(ws(k, v2::oFT) + pq * v2::kPanF);
        const int row = lane >> 1;
        dst[lane] = src[(row < 12 && row >= 4 * pq + 4) ? lane + 50 : lane];
        if (lane < 18) dst[32 + lane] = src[32 + lane];
      }
    }
    __syncwarp();
    const double flag = warp_sum((ng_ != ng_ || nb_ != nb_ || nd_ != nd_ || nm_ != nm_) ? 1.0 : 0.0);
    res[0] = warp_max(ng_ == ng_ ? ng_ : 0.0);
    res[1] = warp_max(nb_ == nb_ ? nb_ : 0.0);
    res[2] = warp_max(nd_ == nd_ ? nd_ : 0.0);
    res[3] = warp_max(nm_ == nm_ ? nm_ : 0.0);
    if (flag > 0.0) res[0] = res[0] + __longlong_as_double(0x7ff8000000000000LL);
    // mu: res_m re-summed in the order of the forward sweep (per row over the stages, then across the rows)
    // (ten loads in flight per trip: the values of the early trips of this sweep have left the L2 -- with four per trip this
    // loop was 2 % of all stall samples, profiles/r2c)
    double smu = 0.0;
    if (lane < 24) {
      for (int k = 0; k < N; k += 10) {
        double v[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) v[j] = k + j < N ? ws_ld(wsc(k + j, v2::oRM)) : 0.0;
#pragma unroll
        for (int j = 0; j < 10; ++j)
          if (k + j < N) smu += v[j];
      }
    }
    mu = warp_sum(smu) / (double)nc_mask;
    __syncwarp();
  }

  __device__ __forceinline__ double shorten(double alpha) const {
    if (alpha < 1.0) {
      if (p.a.alpha_shorten == 0) return alpha * 0.995;
      return alpha * ((1.0 - alpha) * 0.99 + alpha * 0.9999999);
    }
    return alpha;
  }

#ifndef SRBD_K3_PROFILE
#define SRBD_K3_PROFILE 0   // development: per-sweep clock64 sums in bins 48..52 of the iteration histogram
#endif
#if SRBD_K3_PROFILE
  long long prof[5] = {0, 0, 0, 0, 0};
#define SRBD_PROF(i, call) do { const long long t0_ = clock64(); call; prof[i] += clock64() - t0_; } while (0)
#else
#define SRBD_PROF(i, call) call
#endif
  __device__ void solve_one(int qp) {
    set_qp(qp);
    const srbd_ipm_args& a = p.a;
    // ---- d_ocp_qp_init_var (cold start): z = 0, pi = 0, t = max(thr0, -lo), lam = mu0/t (masked rows 0) ------
    int nmask = 0;
    zs = 0;
    const bool lead = kTeam == 0 || wid == 0;   // team mode: warp 0 initialises, runs the recursions and writes the outputs
#pragma unroll 4
    for (int k = lead ? 0 : N + 1; k <= N; ++k) {
      if (lane < 24) wsc(k, v2::oZ)[0] = 0.0;
      if (k < N) {
        if (lane < 12) wsc(k, v2::oPI)[0] = 0.0;
        if (lane < 24) {
          const double lo = __ldg(gDL(k)), mk = mask_of(k);
          double tl = 0.0 - lo;
          tl = a.thr0 > tl ? a.thr0 : tl;
          wsc(k, v2::oT)[0] = tl;
          wsc(k, v2::oLAM)[0] = (a.mu0 / tl) * mk;
          nmask += (mk != 0.0);
          // the masked upper side never moves: t_u = max(thr0, up - v) with up = 0, v = 0; lam_u = 0
          const double tu = a.thr0 > 0.0 ? a.thr0 : 0.0;
          p.sol_t[(size_t)q * N * 48 + k * 48 + 24 + lane] = tu;
          p.sol_lam[(size_t)q * N * 48 + k * 48 + 24 + lane] = 0.0;
        }
      }
    }
    int nc_all = warp_sum_i(nmask);
    if (kTeam) {
      if (wid == 0 && lane == 0) cred[6 * kTeam + 3] = (double)nc_all;
      __threadfence_block();
      __syncthreads();
      nc_all = (int)cred[6 * kTeam + 3];
    }
    // No active row at all (BARRIER_SOFT assembly masks every row): d_ocp_qp_fact_solve_kkt_unconstr, ONE Riccati
    // factorization and solve on the QP itself, iter = 0 (hpipm_d_ocp_qp_kkt.h:54; the reference's own test expects
    // iter == 0, hpipm-cpp/test/ocp_qp_ipm_solver.cpp:56).  It runs through the same call sites as an IPM iteration: at z = 0, pi = 0 the residuals are (g, b) themselves, Gamma = gamma = 0, the full
    // step lands on the solution and the second residual pass evaluates it.
    const bool unc = nc_all == 0;
    const int nc_mask = unc ? 1 : nc_all;
    __syncwarp();
    constexpr bool kFuse = SRBD_K3_FUSE != 0 && !kExp && kTeam == 0;
    double res[4], mu, obj;
    double alpha = 1.0, sp_ = 0.0, sd_ = 0.0;
    int kk = 0;
    if (kExp && p.stat && lead) {
      double* st = p.stat + (size_t)q * p.stat_rows * SRBD_STAT_M;
      for (int i = lane; i < p.stat_rows * SRBD_STAT_M; i += 32) st[i] = 0.0;
      __syncwarp();
    }
    for (;; ++kk) {
      // residuals of the current iterate (kk > 0: the variable update of the previous iteration is fused in)
      if (kFuse) { obj = 0.0; SRBD_PROF(1, sweep_resfac(res, mu, nc_mask, kk > 0, sp_, sd_)); }
      else if (kTeam) SRBD_PROF(0, residuals_team(res, mu, obj, nc_mask, kk > 0, sp_, sd_));
      else SRBD_PROF(0, residuals(res, mu, obj, nc_mask, kk > 0, sp_, sd_));
      if (kExp && p.stat && lead && (!unc || kk > 0)) {   // row kk: mu, the four residual norms, the objective of iterate kk
        if (!unc) stat_row(kk, 5, &mu, 1);
        stat_row(unc ? 0 : kk, 6, res, 4);
        stat_row(unc ? 0 : kk, 10, &obj, 1);
      }
      if (unc ? kk > 0
              : !(kk < a.iter_max && alpha > a.alpha_min &&
                  (res[0] > a.tol_stat || res[1] > a.tol_eq || res[2] > a.tol_ineq || res[3] > a.tol_comp)))
        break;
      if (lead) {
      if (!kFuse) SRBD_PROF(1, sweep_factor());
      // KKT solves of this iteration (one call site per sweep: the sweeps are inlined once).  phase 0: affine /
      // only solve (its backward part was done by the factorization sweep), 1: corrector, 2: conditional centering
      double ap, ad;
      {
        int phase = 0;
        double sigma = 0.0, mua = 0.0, smv = 0.0;
        for (;;) {
          if (phase > 0) SRBD_PROF(2, sweep_backvec(phase, smv));
#if SRBD_K3_PROFILE
          {
            const long long t0_ = clock64();
            sweep_forward(a.pred_corr != 1 || phase > 0 || unc, ap, ad);
            prof[phase > 0 ? 4 : 3] += clock64() - t0_;
          }
#else
          sweep_forward(a.pred_corr != 1 || phase > 0 || unc, ap, ad);
#endif
          if (phase == 0 && !unc) { const double aff = fmin(ap, ad); stat_row(kk + 1, 0, &aff, 1); }
          if (a.pred_corr != 1 || unc) break;
          if (phase == 0) {
            mua = mu_aff(fmin(ap, ad), nc_mask);
            const double tmp = mua / mu;
            sigma = tmp * tmp * tmp;
            smv = sigma * mu;
            smv = smv > a.tau_min ? smv : a.tau_min;
            { const double r2[2] = {mua, sigma}; stat_row(kk + 1, 1, r2, 2); }
            phase = 1;
            continue;
          }
          if (phase == 1 && a.cond_pred_corr == 1) {
            const double muc = mu_aff(fmin(ap, ad), nc_mask);
            if (muc > a.cond_factor * mua) {
              smv = sigma * mu;
              phase = 2;
              continue;
            }
          }
          break;
        }
      }
      if (!a.split_step) {
        const double al = fmin(ap, ad);
        ap = al; ad = al;
      }
      alpha = fmin(ap, ad);
      if (!unc) { const double r2[2] = {ap, ad}; stat_row(kk + 1, 3, r2, 2); }
      sp_ = shorten(ap);
      sd_ = shorten(ad);
      }  // lead
      if (kTeam) {   // the step lengths of the leader reach the team (the barrier also orders its dz, dpi, dt, dlam stores)
        if (wid == 0 && lane == 0) { cred[6 * kTeam] = alpha; cred[6 * kTeam + 1] = sp_; cred[6 * kTeam + 2] = sd_; }
        __threadfence_block();
        __syncthreads();
        alpha = cred[6 * kTeam]; sp_ = cred[6 * kTeam + 1]; sd_ = cred[6 * kTeam + 2];
      }
    }
    if (!lead) return;
    int status;
    const bool nan = (res[0] != res[0]) || (mu != mu);
    if (unc) { kk = 0; status = nan ? 3 : 0; }
    else if (kk == a.iter_max) status = 1;
    else if (alpha <= a.alpha_min) status = 2;
    else if (nan) status = 3;
    else status = 0;
    // ---- outputs (four stages per trip: the loads of a trip are issued before its stores) ----------------------------
    {
      const int l12 = lane < 12 ? lane : 0;
      for (int k0 = 0; k0 <= N; k0 += 4) {
        double vx[4], vp[4], vu[4], vl[4], vt[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + j <= N ? k0 + j : N, nu = k < N ? 12 : 0, kk_ = k < N ? k : N - 1;
          const int l0 = lane < 24 ? lane : 0;   // (zc / pic are per-lane pointers: + min(lane, 23)-ish)
          vx[j] = k == 0 ? __ldg(p.x0 + (size_t)q * 12 + l12) : ws_ld(zc(k, zs) - l0 + nu + l12);
          vp[j] = ws_ld(pic(k > 0 ? k - 1 : 0, zs) - l0 + l12);
          vu[j] = ws_ld(zc(kk_, zs) - l0 + l12);
          vl[j] = ws_ld(wsc(kk_, v2::oLAM));
          vt[j] = ws_ld(wsc(kk_, v2::oT));
          if (unc) { vl[j] = 0.0; vt[j] = 0.0; }  // the unconstrained solve reports lam = t = 0 on the masked rows
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + j;
          if (k <= N) {
            if (lane < 12) {
              p.sol_x[((size_t)q * (N + 1) + k) * 12 + lane] = vx[j];
              if (k > 0) p.sol_pi[((size_t)q * (N + 1) + k) * 12 + lane] = vp[j];
              if (k < N) p.sol_u[((size_t)q * N + k) * 12 + lane] = vu[j];
            }
            if (k < N && lane < 24) {
              p.sol_lam[(size_t)q * N * 48 + k * 48 + lane] = vl[j];
              p.sol_t[(size_t)q * N * 48 + k * 48 + lane] = vt[j];
              if (unc) p.sol_t[(size_t)q * N * 48 + k * 48 + 24 + lane] = 0.0;
            }
          }
        }
      }
    }
    if (kExp && p.ric_P) export_riccati(unc);
    if (lane == 0) {
      p.iter[q] = kk;
      p.status[q] = status;
      for (int i = 0; i < 4; ++i) p.res_max[4 * (size_t)q + i] = res[i];
    }
    __syncwarp();
  }
};

#ifdef SRBD_K3_MAXREG
#define SRBD_K3_BOUNDS __maxnreg__(SRBD_K3_MAXREG)
#else
#define SRBD_K3_BOUNDS __launch_bounds__(32 * v2::kWarps, kTeam ? 1 : v2::kMinCtas)
#endif
// shared-memory doubles of the team's scratch (between the CTA constants and the warp blocks): [kTeam][6] + 4, whole lines
constexpr int kTeamShared = 48;
static_assert(6 * v2::kWarps + 4 <= kTeamShared, "team scratch");
// kSpread: the instantiation for SPARSE batches (at least as many QPs as SMs, fewer than resident warps; see below).  A
// template parameter and not a run-time test because the register allocation of the throughput instantiation has no
// slack: the same early exit compiled into it costs 16 bytes of spills and 4 % (profiles/r2f_spread_ab.txt).
template <int kTma, int kPivot, int kTeam = 0, bool kExp = false, bool kCG = false, bool kSpread = false>
__global__ void SRBD_K3_BOUNDS ipm_srbd_kernel(const SrbdIpmParams p) {
  static_assert(!kSpread || kTeam == 0, "spreading is for the one-warp-per-QP mode");
  static_assert(!kCG || (kTeam == 0 && !kExp && v2::kGP % 2 == 0), "compact BAbt streaming: throughput instantiations, 16-byte aligned panels");
  static_assert(kTeam == 0 || kTeam == v2::kWarps, "a team is the whole CTA");
  extern __shared__ __align__(128) double2 smem2[];  // no static shared memory: the tiles start on 128-byte lines
  if (p.gate && *p.gate != p.gate_value) return;
  if (p.run_gate && *p.run_gate == 0) return;
  if (p.qlist && *p.qcount == 0) return;   // empty rescue list (the usual case)
  double* smem = reinterpret_cast<double*>(smem2);
  int* s_next = reinterpret_cast<int*>(smem + v2::sQ + 16);  // one int per warp behind diag(Q), R
  static_assert(v2::kWarps <= 32, "work-counter slots");
  // CTA-shared constants: Ac and, per constraint row g, the 21 lower-triangle products of its 6-vector
  for (int i = threadIdx.x; i < 288; i += blockDim.x) smem[v2::sAC + i] = p.model->Ac[i];
  if (threadIdx.x < 12) smem[v2::sQ + threadIdx.x] = p.model->m.Q[threadIdx.x];
  if (threadIdx.x == 12) smem[v2::sQ + 12] = p.model->m.R;
  __syncthreads();
  for (int i = threadIdx.x; i < 24 * v2::kW2; i += blockDim.x) {
    const int g = i / v2::kW2, e = i - g * v2::kW2;
    double v = 0.0;
    if (e < 21) {
      int r = 0;
      while ((r + 1) * (r + 2) / 2 <= e) ++r;
      const int c = e - r * (r + 1) / 2, j0 = g < 12 ? 0 : 6;
      v = smem[v2::sAC + g * 12 + j0 + r] * smem[v2::sAC + g * 12 + j0 + c];
    }
    smem[v2::sW + i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Sparse batches (fewer QPs than resident warps; the host launches the full grid of this instantiation): exactly B warps stay,
  // numbered slot-major (warp index first, CTA second), so that every SM runs the same number of solves side by side
  // instead of some SMs carrying two full CTAs and others one (1024 QPs: 6-8 warps on every SM instead of 12 on 23 SMs)
  if (kSpread && !p.qlist && (long long)warp * gridDim.x + blockIdx.x >= p.B) return;
  SrbdSolver<kTma, kPivot, kTeam, kExp, kCG> S(p, smem, smem + v2::kCtaShared + (kTeam ? kTeamShared : 0) + warp * v2::kWarpShared,
                                    kTeam ? blockIdx.x : blockIdx.x * v2::kWarps + warp);
  S.tiles_init();
  if (kCG) S.fill_G_constants(p.gconst);   // the dense record of (QP 0, stage 1): capi.cu takes this instantiation for N >= 2 only
  if (kTeam) {   // one QP per CTA at a time; the helpers see every solve_one call of their leader
    for (;;) {
      __syncthreads();   // (the previous QP's outputs are written, s_next[0] is free)
      if (threadIdx.x == 0) s_next[0] = atomicAdd(p.counter, 1);
      __syncthreads();
      const int idx = s_next[0];
      if (idx >= (p.qlist ? *p.qcount : p.B)) break;
      const int qp = p.qlist ? p.qlist[idx] : idx;
      if (p.frozen && p.frozen[qp]) continue;
      S.solve_one(qp);
      if (threadIdx.x == 0) {
        const int it = p.iter[qp], st = p.status[qp];
        if (st == 1 && p.retry_list) {
          p.retry_list[atomicAdd(p.retry_count, 1)] = qp;
        } else {
          atomicAdd((unsigned long long*)&p.bstats->solves, 1ull);
          atomicAdd((unsigned long long*)&p.bstats->iter_sum, (unsigned long long)it);
          atomicAdd((unsigned long long*)&p.bstats->status_count[st < 0 || st > 4 ? 4 : st], 1ull);
          atomicAdd((unsigned long long*)&p.bstats->iter_hist[it < SRBD_HIST_BINS ? it : SRBD_HIST_BINS - 1], 1ull);
          for (int i = 0; i < 4; ++i)
            atomicMax((unsigned long long*)&p.bstats->res_max[i],
                      (unsigned long long)__double_as_longlong(p.res_max[4 * (size_t)qp + i]));
        }
      }
    }
#if SRBD_K3_PROFILE
    if (threadIdx.x == 0)
      for (int i = 0; i < 5; ++i) atomicAdd((unsigned long long*)&p.bstats->iter_hist[48 + i], (unsigned long long)S.prof[i]);
#endif
    return;
  }
  long long it_sum = 0, solves = 0;
  int st_cnt[5] = {0, 0, 0, 0, 0};
  double rmax[4] = {0.0, 0.0, 0.0, 0.0};
  for (;;) {
    if (lane == 0) s_next[warp] = atomicAdd(p.counter, 1);
    __syncwarp();
    const int idx = s_next[warp];
    __syncwarp();
    if (idx >= (p.qlist ? *p.qcount : p.B)) break;
    const int qp = p.qlist ? p.qlist[idx] : idx;
    if (p.frozen && p.frozen[qp]) continue;
    S.solve_one(qp);
    if (lane == 0) {
      const int it = p.iter[qp], st = p.status[qp];
      if (st == 1 && p.retry_list) {  // iter_max only: see capi.cu, solve_srbd_variant
        p.retry_list[atomicAdd(p.retry_count, 1)] = qp;   // counted by the rescue pass
      } else {
        it_sum += it;
        solves += 1;
        st_cnt[st < 0 || st > 4 ? 4 : st] += 1;
        atomicAdd((unsigned long long*)&p.bstats->iter_hist[it < SRBD_HIST_BINS ? it : SRBD_HIST_BINS - 1], 1ull);
        for (int i = 0; i < 4; ++i) rmax[i] = fmax(rmax[i], p.res_max[4 * (size_t)qp + i]);
      }
    }
  }
#if SRBD_K3_PROFILE
  if (lane == 0)
    for (int i = 0; i < 5; ++i) atomicAdd((unsigned long long*)&p.bstats->iter_hist[48 + i], (unsigned long long)S.prof[i]);
#endif
  if (lane == 0 && solves > 0) {
    atomicAdd((unsigned long long*)&p.bstats->solves, (unsigned long long)solves);
    atomicAdd((unsigned long long*)&p.bstats->iter_sum, (unsigned long long)it_sum);
    for (int i = 0; i < 5; ++i)
      if (st_cnt[i]) atomicAdd((unsigned long long*)&p.bstats->status_count[i], (unsigned long long)st_cnt[i]);
    for (int i = 0; i < 4; ++i)
      atomicMax((unsigned long long*)&p.bstats->res_max[i], (unsigned long long)__double_as_longlong(rmax[i]));
  }
}

}  // namespace srbd
