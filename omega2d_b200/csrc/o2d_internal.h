// o2d_internal.h -- host/device shared declarations of the omega2d_b200 library (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace o2d {

enum : int { CORE_RM = 0, CORE_EXP = 1, CORE_WL = 2, CORE_V2 = 3, CORE_V3 = 4 };
// what a pair-sum kernel accumulates (component lists in o2d_pp_impl.cuh)
enum : int { RES_U = 0, RES_UW = 1, RES_UE = 2, RES_UG = 3, RES_VORT = 4 };

constexpr int kTile = 1024;     // the packed source array is padded to a multiple of this (also the points-on-panels tile)
constexpr int kBigTile = 512;   // sources per shared-memory stage of the pair-sum kernels (8 KiB)
constexpr int kStages = 4;      // stages in flight per CTA (2 ... 16 measured the same)
constexpr int kThreads = 256;   // threads per CTA of the pair-sum kernels
constexpr int kMaxOut = 6;
// "fine" configuration of the same kernel for small problems: a CTA-tile is 128*R x 256 pairs instead of
// 256*R x 1024, so a few thousand particles still spread over all SMs and no CTA runs longer than a few us
constexpr int kFineTile = 256;
constexpr int kFineThreads = 128;

struct OutPtrs {
  float* p[kMaxOut];
};

// arguments of one pair-sum launch (passed by value in kernel parameter space)
//
// Work of a launch = the (target block, source tile) steps in block-major order, nblocks * ntiles of them.
// CTA k owns the contiguous run of steps [k*steps_per_cta, (k+1)*steps_per_cta): it walks the tiles of a target
// block, and when the run crosses into the next block it stores what it has and reloads its targets.  A block
// that lies inside one run is added straight into the outputs; a block cut between runs leaves its pieces in
// `partial` (at most two per CTA: the head and the tail of its run) and pp_combine_kernel sums them in CTA order,
// so the result does not depend on scheduling.
struct PPArgs {
  const float4* packed;      // packed source pairs: 2 float4 per pair, TILE float4 per tile
  int ntiles;                // tiles of the packed source array
  int nblocks;               // target blocks of THREADS*R targets
  long long steps_per_cta;   // S: steps per CTA (the last CTA may own fewer)
  const float* tx;
  const float* ty;
  const float* tr;           // nullptr: inert targets
  int nt;
  OutPtrs out;
  double* partial;           // [cta][2][component][THREADS*R]
};

struct PanelArgs {
  const float4* packed;   // kPanelRecF4 float4 per panel, 2*kPanelRecF4 per panel PAIR (layout: pack_panels_kernel)
  int npanels_padded;     // multiple of kPanelTile (what is packed: zero-strength padding behind the real panels)
  int npanels_even;       // the real panels rounded up to a pair: what the kernel walks (the last tile may be partial)
  int tiles_per_chunk;
  const float* tx;
  const float* ty;
  int nt;
  OutPtrs out;
  double* partial;
};

constexpr int kPanelTile = 256;  // panels (or particles, for points-on-panels) per smem tile
#ifndef O2D_PAN_BIG_R
#define O2D_PAN_BIG_R 2
#endif
constexpr int kPanBigR = O2D_PAN_BIG_R;   // targets per thread of the panels-on-points kernel when there are enough targets
constexpr int kPanelRecF4 = 4;   // float4 per packed panel (16 floats: node0, H, four complex series coefficients, thresholds)

// number of raw accumulated components per result kind (mirrors ResTraits in o2d_pp_impl.cuh)
inline int res_ncomp(int res) {
  switch (res) { case 0: return 2; case 1: return 4; case 2: return 3; case 3: return 6; case 4: return 1; }
  return 0;
}
#ifndef O2D_UW_R
#define O2D_UW_R 3
#endif
#ifndef O2D_UW_UNROLL
#define O2D_UW_UNROLL 1
#endif
#ifndef O2D_UE_R
#define O2D_UE_R 4
#endif
#ifndef O2D_UE_WARPS
#define O2D_UE_WARPS 16
#endif
#ifndef O2D_UE_UNROLL
#define O2D_UE_UNROLL 4
#endif
// ---- pair-sum kernel shape per (core, result kind, configuration) --------------------------------------------
// targets per thread R, resident warps per SM the register allocation must allow, and inner-loop unroll (source
// pairs).  Measured on B200, 1M x 1M (pairs/clk/SM; profiles/r02_pp_shape_search.md): RM and V2 velocity are
// fastest with 16 resident warps of 128 registers and R = 6 (RM 12.89, V2 12.90; 24 warps of 80 registers with
// R = 4 give 12.56 / 12.56); WL wants R = 4 (10.75 against 10.37), EXP and V3 24 warps and R = 4 (EXP 11.04
// against 10.33; V3 with uniform radii 7.53 against 6.91); the gradient-type kernels (128 registers, R = 3) lose
// to any unrolling (RM velocity+gradient 7.23 / 7.22 / 6.98 / 6.84 at 1 / 2 / 8 / 16); velocity + shear rate: R = 4,
// 16 warps, unroll 4 (RM 9.56, V2 uniform 9.05; the 24-warp R = 2 shape of the first run-walking version 9.37 / 8.59),
// except EXP and V3, which keep R = 2 / 24 warps / unroll 8 there as well (EXP 7.61 against 7.12, V3 4.58 against 4.38).
// EXP takes 24 warps x 2 targets for velocity+vorticity and velocity+gradient too (7.17 against 6.76, 5.94 against 5.59;
// V3: 4.50 / 4.54 and 3.74 / 3.76, unchanged, stays at 16 warps x 3).
constexpr int pp_targets(int core, int res, bool fine) {
  if (fine) return (res == RES_U || res == RES_VORT) ? 2 : 1;
  switch (res) {
    case RES_U:    return (core == CORE_WL || core == CORE_EXP || core == CORE_V3) ? 4 : 6;
    case RES_VORT: return 4;
    case RES_UE:   return (core == CORE_EXP || core == CORE_V3) ? 2 : O2D_UE_R;
    case RES_UW:   return core == CORE_EXP ? 2 : O2D_UW_R;
    default:       return core == CORE_EXP ? 2 : 3;   // RES_UG
  }
}
constexpr int pp_min_warps(int core, int res, bool fine) {
  if (res == RES_UE) return (core == CORE_EXP || core == CORE_V3) ? 24 : O2D_UE_WARPS;
  if (res == RES_VORT) return 24;
  if (core == CORE_EXP) return 24;   // every result kind: the far/near votes want resident warps more than registers
  if (res != RES_U) return 16;
  return (fine || core == CORE_EXP || core == CORE_V3) ? 24 : 16;
}
constexpr int pp_unroll(int core, int res, bool fine) {
  if (fine || core == CORE_EXP || core == CORE_WL) return 4;
  if (res == RES_UE) return (core == CORE_EXP || core == CORE_V3) ? 8 : O2D_UE_UNROLL;
  if (res == RES_VORT) return 8;
  if (res == RES_UW) return O2D_UW_UNROLL;
  if (res == RES_U) return core == CORE_V3 ? 8 : 16;
  return 1;
}
// threads per CTA.  Coarse configuration: the 16-warp kernels run as ONE 512-thread CTA per SM with one equal run
// each; the 24-warp kernels as three 256-thread CTAs with short runs (measured as one 768-thread CTA: EXP 10.5
// against 11.0 pairs/clk/SM, V3 7.2 against 7.3).  The V2 core also keeps 256-thread CTAs (two per SM, short runs):
// equal at 1M particles, but 6 % faster than one 512-thread CTA at 100k-400k (V2 uniform, 200k: 10.88 vs 11.52 ms).
// See pp_plan (o2d_api.cu).
constexpr int pp_threads(int core, int res, bool fine) {
  return fine ? kFineThreads : ((pp_min_warps(core, res, false) == 16 && core != CORE_V2) ? 512 : kThreads);
}
inline int res_targets_per_thread(int core, int res, bool fine = false) { return pp_targets(res == RES_VORT ? CORE_RM : core, res, fine); }
// targets per block of the (core, res) kernel
inline int pp_block_targets(int core, int res, bool fine) {
  const int c = res == RES_VORT ? CORE_RM : core;
  return pp_threads(c, res, fine) * pp_targets(c, res, fine);
}

// ---- launchers implemented per translation unit ------------------------------------------------
// pair-sum kernel for (core, res); grid = nctas CTAs, each a run of a.steps_per_cta steps
// fine: tiles in `a` are kFineTile sources and a CTA holds kFineThreads threads
cudaError_t launch_pp(int core, int res, const PPArgs& a, int nctas, cudaStream_t st, bool fine = false);
// sums the pieces of the blocks that were cut between CTAs and folds them into a.out (a = the launch it follows)
cudaError_t launch_pp_combine(int core, int res, const PPArgs& a, cudaStream_t st, bool fine = false);
// resident CTAs per SM of the (core,res) kernel; cached
int pp_occupancy(int core, int res, bool fine = false);

// source packing: SoA (x,y,s,r) -> pairwise {-x0,-x1,-y0,-y1}{s0,s1,rho0,rho1}; writes ns_padded (even) entries
// for sources j0, j0+1, ...: real ones below ns, zero-strength copies of the last particle from ns on.  mode: core
// id, or 100 for the Gaussian-vorticity terms, or 101 for "strength only" (points-on-panels sources).
cudaError_t launch_pack_sources(int mode, int ns, const float* sx, const float* sy, const float* ss,
                                const float* sr, float4* packed, int ns_padded, cudaStream_t st, int j0 = 0);

// panels
// writes np_padded (even) packed panels for panels j0, j0+1, ...: real ones below np, zero-strength copies of the last from np on
cudaError_t launch_pack_panels(int np, const float* nx, const float* ny, const uint32_t* idx,
                               const float* vs, const float* ss, float4* packed, int np_padded,
                               cudaStream_t st, int j0 = 0);
// rpt: targets per thread (1 or kPanBigR); grid.x covers 128*rpt targets per CTA
cudaError_t launch_panels_on_points(const PanelArgs& a, int nchunks, int rpt, cudaStream_t st);
cudaError_t launch_panels_combine(const double* partial, int nchunks, int nt, const OutPtrs& out, cudaStream_t st);
// points-on-panels: targets are panels (one per thread), sources are packed particles
struct PtPanArgs {
  const float4* packed_src;   // particle pairs, as for the pp kernels (mode 101)
  int ntiles;                 // kTile-sized tiles of particles
  int tiles_per_chunk;
  const float* nx; const float* ny; const uint32_t* idx;
  const float* area;
  int np;
  float* pu; float* pv;
  double* partial;            // [chunk][2][np] or nullptr
};
cudaError_t launch_points_on_panels(const PtPanArgs& a, int nchunks, cudaStream_t st);
cudaError_t launch_ptpan_combine_geom(const double* partial, int nchunks, const PtPanArgs& a, cudaStream_t st);
// out[2][np] = sum over chunks of partial[chunk][2][np] (zeros when nchunks == 0)
cudaError_t launch_ptpan_reduce_chunks(const double* partial, int nchunks, int np, double* out, cudaStream_t st);
int panels_occupancy(int rpt);
int ptpan_occupancy();
int ptpan_block_panels();   // panels per CTA of points_on_panels_kernel

// elementwise helpers
cudaError_t launch_zero(int n, float* u, float* v, float* w, float* e, cudaStream_t st);
cudaError_t launch_finalize(int n, float* u, float* v, float* w, float* e, double fsx, double fsy,
                            cudaStream_t st);

// FFMA / FFMA2 throughput micro-kernel; returns elapsed ms for `iters` iterations
cudaError_t run_fp32_microbench(int packed, int blocks, int threads, int iters, float* sink,
                                cudaStream_t st);
constexpr int kMicrobenchFmaPerIter = 64;  // lane-FMAs per thread per iteration

}  // namespace o2d
