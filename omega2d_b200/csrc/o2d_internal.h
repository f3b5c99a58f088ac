// o2d_internal.h -- host/device shared declarations of the omega2d_b200 library (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace o2d {

enum : int { CORE_RM = 0, CORE_EXP = 1, CORE_WL = 2, CORE_V2 = 3, CORE_V3 = 4 };
// what a pair-sum kernel accumulates (component lists in o2d_pp_impl.cuh)
enum : int { RES_U = 0, RES_UW = 1, RES_UE = 2, RES_UG = 3, RES_VORT = 4 };

constexpr int kTile = 1024;     // sources per shared-memory tile (16 KiB per stage)
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
struct PPArgs {
  const float4* packed;   // packed source pairs: 2 float4 per pair, kTile float4 per tile
  int tile_begin;         // logical tile range [tile_begin, tile_end) of this launch
  int tile_end;
  int tiles_per_chunk;    // blockIdx.y selects the chunk inside the range
  int skip_begin;         // logical tiles >= skip_begin map to physical tile + skip_len
  int skip_len;           //   (multi-GPU: hop over the rank's own block in the gathered array)
  int chunk_offset;       // first chunk index of this launch inside the partials buffer
  const float* tx;
  const float* ty;
  const float* tr;        // nullptr: inert targets
  int nt;
  OutPtrs out;
  double* partial;        // nullptr: single chunk, fold straight into `out`
};

struct PanelArgs {
  const float4* packed;   // 4 float4 per panel PAIR (layout: pack_panels_kernel)
  int npanels_padded;     // multiple of kPanelTile
  int tiles_per_chunk;
  const float* tx;
  const float* ty;
  int nt;
  OutPtrs out;
  double* partial;
};

constexpr int kPanelTile = 256;  // panels (or particles, for points-on-panels) per smem tile

// number of raw accumulated components per result kind (mirrors ResTraits in o2d_pp_impl.cuh)
inline int res_ncomp(int res) {
  switch (res) { case 0: return 2; case 1: return 4; case 2: return 3; case 3: return 6; case 4: return 1; }
  return 0;
}
inline int res_targets_per_thread(int res, bool fine = false) {
  const int r = (res == 0 || res == 4) ? 4 : (res == 2 ? 2 : 3);   // mirrors ResTraits<RES>::R
  return fine ? r / 2 : r;
}

// ---- launchers implemented per translation unit ------------------------------------------------
// pair-sum kernel for (core, res); grid = (target blocks, nchunks)
// fine: tile indices in `a` count kFineTile-source tiles and a CTA holds kFineThreads threads
cudaError_t launch_pp(int core, int res, const PPArgs& a, int nchunks, cudaStream_t st, bool fine = false);
cudaError_t launch_pp_combine(int res, const double* partial, int nchunks, int nt, const OutPtrs& out,
                              cudaStream_t st);
// resident CTAs per SM of the (core,res) kernel; cached
int pp_occupancy(int core, int res, bool fine = false);

// source packing: SoA (x,y,s,r) -> pairwise {-x0,-x1,-y0,-y1}{s0,s1,rho0,rho1}, padded to a whole
// number of tiles with zero-strength copies of the last particle.  mode: core id, or 100 for
// the Gaussian-vorticity terms, or 101 for "strength only" (points-on-panels sources).
cudaError_t launch_pack_sources(int mode, int ns, const float* sx, const float* sy, const float* ss,
                                const float* sr, float4* packed, int ns_padded, cudaStream_t st);

// panels
cudaError_t launch_pack_panels(int np, const float* nx, const float* ny, const uint32_t* idx,
                               const float* vs, const float* ss, float4* packed, int np_padded,
                               cudaStream_t st);
// rpt: targets per thread (1 or 2); grid.x covers 128*rpt targets per CTA
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

// elementwise helpers
cudaError_t launch_zero(int n, float* u, float* v, float* w, float* e, cudaStream_t st);
cudaError_t launch_finalize(int n, float* u, float* v, float* w, float* e, double fsx, double fsy,
                            cudaStream_t st);

// FFMA / FFMA2 throughput micro-kernel; returns elapsed ms for `iters` iterations
cudaError_t run_fp32_microbench(int packed, int blocks, int threads, int iters, float* sink,
                                cudaStream_t st);
constexpr int kMicrobenchFmaPerIter = 64;  // lane-FMAs per thread per iteration

}  // namespace o2d
