// o2d_pp_impl.cuh -- the particles-on-points pair-sum kernel (one template, instantiated per core).
//
// Replaces the reference loop nests of points_affect_points (reference src/Influence.h:197-249
// inert targets, :347-403 blob targets) and points_affect_points_vorticity
// (src/InfluenceVort.h:106-118).
//
// Work decomposition
//   grid.x = target blocks of THREADS*R targets, grid.y = source chunks of `tiles_per_chunk` tiles.
//   A thread owns R targets in registers for its whole life and walks its chunk of the packed
//   source array tile by tile.  Source tiles (TILE sources = TILE*16 bytes) are staged into
//   shared memory by ONE TMA bulk copy per tile (cp.async.bulk + mbarrier), double buffered, and
//   read back with warp-broadcast LDS.128.
// Arithmetic
//   Sources are stored pairwise -- {x0,x1,y0,y1} {s0,s1,rho0,rho1} -- so that one LDS.128 feeds
//   both lanes of Blackwell's packed FP32 instructions (FADD2/FMUL2/FFMA2): every FP32
//   instruction in the inner loop evaluates two source-target pairs.
// Accumulation (SURVEY 7 "hard parts": a single FP32 accumulator fails the 1e-5 tolerance)
//   per tile: float partial sums (two lanes x TILE/2 terms), folded into float64 once per tile;
//   the float64 sums are either added straight into the float outputs ("tu[i] += accum", as the
//   reference) when the launch has one source chunk, or written to a partials buffer that
//   o2d_pp_combine sums in fixed chunk order -- results are bit-reproducible run to run.
#pragma once
#include "o2d_device.cuh"
#include "o2d_internal.h"

namespace o2d {

// raw accumulated components per result kind
//   RES_U    : U = sum r2 dy, V = sum r2 dx                                  (r2 = ss*K)
//   RES_UW   : U, V, W = sum bb d2, R = sum r2                                (bb = ss*bbb)
//   RES_UE   : U, V, E = sum |bb| d2
//   RES_UG   : U, V, XY = sum bb dx dy, XX = sum bb dx^2, YY = sum bb dy^2, R
//   RES_VORT : W = sum c exp2(-d2 a)      (Gaussian vorticity; a = log2e/sr^2, c = 2 ss/sr^2)
// how they map onto the reference's per-pair updates is in apply_outputs() below.
template <int RES> struct ResTraits;
template <> struct ResTraits<RES_U>    { static constexpr int NCOMP = 2; static constexpr int R = 4; static constexpr bool GRAD = false; };
template <> struct ResTraits<RES_UW>   { static constexpr int NCOMP = 4; static constexpr int R = 3; static constexpr bool GRAD = true; };
template <> struct ResTraits<RES_UE>   { static constexpr int NCOMP = 3; static constexpr int R = 2; static constexpr bool GRAD = true; };
template <> struct ResTraits<RES_UG>   { static constexpr int NCOMP = 6; static constexpr int R = 3; static constexpr bool GRAD = true; };
template <> struct ResTraits<RES_VORT> { static constexpr int NCOMP = 1; static constexpr int R = 4; static constexpr bool GRAD = false; };

// Fold the float64 component sums of one target into the float outputs exactly like the
// reference's epilogue "tu[0][i] += accumu" (float += double; src/Influence.h:359-360,378-380).
//   kernelu_0v_0b   src/Kernels.h:90-91    tu -= r2 dy ; tv += r2 dx
//   kerneluw_0v_0b  src/Kernels.h:176-178  tw += (bb dx^2 + r2) - (-bb dy^2 - r2) = bb d2 + 2 r2
//   kernelue_0v_0b  src/Kernels.h:225-230  te += sqrt(dudx^2 + asym^2/4) = |bb| d2 / 2  (identity:
//                   (dx dy)^2 + (dx^2-dy^2)^2/4 = (dx^2+dy^2)^2/4)
//   kernelug_0v_0b  src/Kernels.h:129-132  tux += -bb dx dy ; tuy += -bb dy^2 - r2 ;
//                                          tvx += bb dx^2 + r2 ; tvy += bb dx dy
template <int RES>
__device__ __forceinline__ void apply_outputs(const OutPtrs& o, int i, const double* c) {
  if (RES == RES_VORT) {
    o.p[0][i] = (float)((double)o.p[0][i] + c[0]);
    return;
  }
  o.p[0][i] = (float)((double)o.p[0][i] - c[0]);
  o.p[1][i] = (float)((double)o.p[1][i] + c[1]);
  if (RES == RES_UW) {
    o.p[2][i] = (float)((double)o.p[2][i] + (c[2] + 2.0 * c[3]));
  } else if (RES == RES_UE) {
    o.p[2][i] = (float)((double)o.p[2][i] + 0.5 * c[2]);
  } else if (RES == RES_UG) {
    o.p[2][i] = (float)((double)o.p[2][i] - c[2]);            // tux
    o.p[3][i] = (float)((double)o.p[3][i] + (c[3] + c[5]));   // tvx
    o.p[4][i] = (float)((double)o.p[4][i] - (c[4] + c[5]));   // tuy
    o.p[5][i] = (float)((double)o.p[5][i] + c[2]);            // tvy
  }
}

// factor applied to component c when a tile's float partial sums are folded into float64: the deferred -2 of the
// gradient factor for the components built from bb (|bb| for the shear rate), 1 otherwise
template <int CORE, int RES>
__device__ __forceinline__ constexpr double bb_fold_scale(int c) {
  if (!CoreTraits<CORE>::kBbDeferred) return 1.0;
  if (RES == RES_UW && c == 2) return -2.0;
  if (RES == RES_UE && c == 2) return 2.0;
  if (RES == RES_UG && (c == 2 || c == 3 || c == 4)) return -2.0;
  return 1.0;
}

// one source pair's terms onto one target's float partial sums (component lists: ResTraits above)
// DEFER: Bb is the deferred gradient factor of core_eval (bb = r2 * Bb, the -2 applied at the float64 fold)
template <int RES, bool DEFER>
__device__ __forceinline__ void pair_accumulate(float2 (&part)[ResTraits<RES>::NCOMP], float2 str, float2 K, float2 Bb,
                                                float2 dx, float2 dy, float2 d2) {
  const float2 r2 = mul2(str, K);
  part[0] = fma2(r2, dy, part[0]);
  part[1] = fma2(r2, dx, part[1]);
  if (RES == RES_UW) {
    const float2 bb = mul2(DEFER ? r2 : str, Bb);
    part[2] = fma2(bb, d2, part[2]);
    part[3] = add2(part[3], r2);
  } else if (RES == RES_UE) {
    const float2 bb = mul2(DEFER ? r2 : str, Bb);
    const float2 ab = make_float2(fabsf(bb.x), fabsf(bb.y));
    part[2] = fma2(ab, d2, part[2]);
  } else if (RES == RES_UG) {
    const float2 bb = mul2(DEFER ? r2 : str, Bb);
    const float2 bx = mul2(bb, dx);
    const float2 by = mul2(bb, dy);
    part[2] = fma2(bx, dy, part[2]);
    part[3] = fma2(bx, dx, part[3]);
    part[4] = fma2(by, dy, part[4]);
    part[5] = add2(part[5], r2);
  }
}

// Sum one staged tile of TILE sources onto the thread's R targets (float partial sums, two lanes).
// UNI = all R targets share one radius term (see core_eval): rho+tau is formed once per source pair.
// rho_max (EXP core only): the largest source radius term of this tile, for the far test of core_eval.
template <int CORE, int RES, int TILE, int R, int UNROLL, bool UNI>
__device__ __forceinline__ void tile_sum(const float4* __restrict__ s, const float2 (&tx)[R], const float2 (&ty)[R],
                                         const float2 (&tau)[R], float2 (&part)[R][ResTraits<RES>::NCOMP],
                                         const float rho_max) {
  using T = ResTraits<RES>;
  // EXP: beyond d2 > lim[k] every pair of this tile with target k is in the reference's "d2/sig2 > 16" branch
  float lim[R];
#pragma unroll
  for (int k = 0; k < R; ++k) lim[k] = (T::GRAD ? 24.0f : 16.0f) * (rho_max + tau[k].x);
#pragma unroll UNROLL
  for (int p = 0; p < TILE / 2; ++p) {
    const float4 A = s[2 * p];      // x0 x1 y0 y1   (stored NEGATED: dx = tx + (-sx))
    const float4 B = s[2 * p + 1];  // s0 s1 rho0 rho1
    const float2 nsx = f2(A.x, A.y), nsy = f2(A.z, A.w);
    const float2 str = f2(B.x, B.y);
    const float2 rho = (UNI && RES != RES_VORT) ? add2(f2(B.z, B.w), tau[0]) : f2(B.z, B.w);
    if constexpr (CORE == CORE_EXP && RES != RES_VORT) {
      // compact-exponential core: ONE warp vote per source pair decides whether all R targets of all lanes are in
      // the reference's far branch (K = 1/d2); only the few source pairs that have a close target anywhere in the
      // warp go through core_eval's per-target test and general path
      float2 dxk[R], dyk[R], d2k[R];
      bool allfar = true;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        dxk[k] = add2(tx[k], nsx);
        dyk[k] = add2(ty[k], nsy);
        d2k[k] = fma2(dyk[k], dyk[k], mul2(dxk[k], dxk[k]));
        allfar = allfar && (d2k[k].x > lim[k]) && (d2k[k].y > lim[k]);
      }
      if (__all_sync(0xffffffffu, allfar)) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const float2 K = rcp2(d2k[k]);
          const float2 Bb = T::GRAD ? mul2(mul2(f2(-2.0f), K), K) : f2(0.0f);
          pair_accumulate<RES, false>(part[k], str, K, Bb, dxk[k], dyk[k], d2k[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < R; ++k) {
          float2 K, Bb, d2;
          core_eval<CORE, T::GRAD, UNI>(dxk[k], dyk[k], rho, tau[k], K, Bb, d2, lim[k]);
          pair_accumulate<RES, false>(part[k], str, K, Bb, dxk[k], dyk[k], d2);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float2 dx = add2(tx[k], nsx);
        const float2 dy = add2(ty[k], nsy);
        if (RES == RES_VORT) {
          // src/InfluenceVort.h:109-114; str = 2 ss/sr^2, rho = log2(e)/sr^2
          const float2 d2 = fma2(dy, dy, mul2(dx, dx));
          const float2 arg = mul2(d2, rho);
          const float2 e = make_float2(mufu_ex2(-arg.x), mufu_ex2(-arg.y));
          part[k][0] = fma2(str, e, part[k][0]);
        } else {
          float2 K, Bb, d2;
          core_eval<CORE, T::GRAD, UNI, RES != RES_UG>(dx, dy, rho, tau[k], K, Bb, d2, lim[k]);
          pair_accumulate<RES, CoreTraits<CORE>::kBbDeferred>(part[k], str, K, Bb, dx, dy, d2);
        }
      }
    }
  }
}

template <int CORE, int RES, int THREADS, int TILE, int R = ResTraits<RES>::R, int UNROLL = 4>
__global__ void __launch_bounds__(THREADS)
pp_kernel(const PPArgs a) {
  using T = ResTraits<RES>;
  constexpr int NC = T::NCOMP;
  constexpr uint32_t TILE_BYTES = TILE * 16;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* stage0 = reinterpret_cast<float4*>(smem_raw);
  float4* stage1 = reinterpret_cast<float4*>(smem_raw + TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * TILE_BYTES);
  __shared__ float wmax[2 * (THREADS / 32)];   // EXP core: per-warp maxima of the tile's radius terms, double buffered

  const int tid = threadIdx.x;
  const int tile_begin = a.tile_begin + blockIdx.y * a.tiles_per_chunk;
  const int tile_end = min(tile_begin + a.tiles_per_chunk, a.tile_end);
  const int ntiles = tile_end - tile_begin;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  // logical -> physical tile (TILE float4 per tile: 2 per source pair)
  auto tile_ptr = [&](int lt) -> const float4* {
    const int pt = lt >= a.skip_begin ? lt + a.skip_len : lt;
    return a.packed + (size_t)pt * TILE;
  };
  if (tid == 0 && ntiles > 0) {
    mbar_arrive_expect_tx(&bars[0], TILE_BYTES);
    tma_load_1d(stage0, tile_ptr(tile_begin), TILE_BYTES, &bars[0]);
  }

  // ---- this thread's targets ------------------------------------------------------------------
  const int tbase = blockIdx.x * (THREADS * R) + tid;
  float2 tx[R], ty[R], tau[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int i = tbase + k * THREADS;
    const bool ok = i < a.nt;
    tx[k] = f2(ok ? a.tx[i] : 0.0f);
    ty[k] = f2(ok ? a.ty[i] : 0.0f);
    float t = 0.0f;
    if (RES != RES_VORT && a.tr != nullptr && ok) t = radius_term<CORE>(a.tr[i]);
    tau[k] = f2(t);
  }
  // every thread's R targets share one radius term (inert targets; uniform-radius particle sets,
  // the reference's normal case): take the cheaper tile_sum<UNI>.  The flag is made CTA-uniform so
  // the tile loop's barriers stay convergent.  Out-of-range targets carry 0: they only send the
  // last CTA of a blob-target launch down the general path.
  bool uni_t = true;
#pragma unroll
  for (int k = 1; k < R; ++k) uni_t = uni_t && (tau[k].x == tau[0].x);
  const bool uni = __syncthreads_and(uni_t) != 0;

  double acc[R][NC];
#pragma unroll
  for (int k = 0; k < R; ++k)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[k][c] = 0.0;

  // ---- walk the source tiles --------------------------------------------------------------------
  for (int t = 0; t < ntiles; ++t) {
    // every thread is done with tile t-1, so its stage may be overwritten by the prefetch of t+1
    __syncthreads();
    if (tid == 0 && t + 1 < ntiles) {
      uint64_t* bar = &bars[(t + 1) & 1];
      mbar_arrive_expect_tx(bar, TILE_BYTES);
      tma_load_1d((t + 1) & 1 ? stage1 : stage0, tile_ptr(tile_begin + t + 1), TILE_BYTES, bar);
    }
    mbar_wait(&bars[t & 1], (uint32_t)((t >> 1) & 1));
    const float4* s = (t & 1) ? stage1 : stage0;

    // EXP core: largest source radius term of the tile (CTA-wide max of the rho lanes; radius terms are >= 0)
    float rho_max = 0.0f;
    if (CORE == CORE_EXP && RES != RES_VORT) {
      float m = 0.0f;
      for (int p = tid; p < TILE / 2; p += THREADS) { const float4 B = s[2 * p + 1]; m = fmaxf(m, fmaxf(B.z, B.w)); }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      if ((tid & 31) == 0) wmax[(t & 1) * (THREADS / 32) + (tid >> 5)] = m;
      __syncthreads();
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) rho_max = fmaxf(rho_max, wmax[(t & 1) * (THREADS / 32) + w]);
    }

    float2 part[R][NC];
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
      for (int c = 0; c < NC; ++c) part[k][c] = f2(0.0f);

    if (uni) tile_sum<CORE, RES, TILE, R, UNROLL, true>(s, tx, ty, tau, part, rho_max);
    else     tile_sum<CORE, RES, TILE, R, UNROLL, false>(s, tx, ty, tau, part, rho_max);
    // fold this tile's float partial sums into the float64 accumulators
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double sum = (double)part[k][c].x + (double)part[k][c].y;
        acc[k][c] += bb_fold_scale<CORE, RES>(c) * sum;
      }
  }

  // ---- epilogue -----------------------------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int i = tbase + k * THREADS;
    if (i >= a.nt) continue;
    if (a.partial == nullptr) {
      apply_outputs<RES>(a.out, i, acc[k]);
    } else {
      // partial[(chunk*NC + c) * nt + i], coalesced over i
      double* dst = a.partial + ((size_t)(a.chunk_offset + blockIdx.y) * NC) * (size_t)a.nt + i;
#pragma unroll
      for (int c = 0; c < NC; ++c) dst[(size_t)c * a.nt] = acc[k][c];
    }
  }
}

// sum the per-chunk float64 partials in fixed chunk order, then fold into the outputs
template <int RES>
__global__ void __launch_bounds__(256)
pp_combine_kernel(const double* __restrict__ partial, int nchunks, int nt, const OutPtrs out) {
  constexpr int NC = ResTraits<RES>::NCOMP;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  double c[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) c[k] = 0.0;
  for (int ch = 0; ch < nchunks; ++ch) {
    const double* src = partial + ((size_t)ch * NC) * (size_t)nt + i;
#pragma unroll
    for (int k = 0; k < NC; ++k) c[k] += src[(size_t)k * nt];
  }
  apply_outputs<RES>(out, i, c);
}

}  // namespace o2d

// ---- per-core launcher (explicitly instantiated in o2d_pp_<core>.cu) ---------------------------
namespace o2d {

template <int CORE, int RES, bool FINE>
struct PPConfig {
  static constexpr int THREADS = FINE ? kFineThreads : kThreads;
  static constexpr int TILE = FINE ? kFineTile : kTile;
  static constexpr int R = FINE ? ResTraits<RES>::R / 2 : ResTraits<RES>::R;
  static constexpr size_t SMEM = 2 * (size_t)TILE * 16 + 64;
  // inner-loop unroll: 8 source pairs for the big configuration (4 for the branchy EXP core), 4 for the small tiles
  static constexpr int UNROLL = (FINE || CORE == CORE_EXP) ? 4 : 8;
  static auto kernel() { return pp_kernel<CORE, RES, THREADS, TILE, R, UNROLL>; }
};

template <int CORE, int RES, bool FINE>
static cudaError_t launch_one(const PPArgs& a, int nchunks, cudaStream_t st) {
  using C = PPConfig<CORE, RES, FINE>;
  auto kern = C::kernel();
  // two 16 KiB stages + barriers: under the 48 KiB a kernel gets without opting in, so there is no per-device
  // function attribute to set (a process may drive several devices, o2d_group.cu)
  static_assert(C::SMEM <= 48 * 1024, "dynamic shared memory above the default limit needs cudaFuncSetAttribute per device");
  const int nb = (a.nt + C::THREADS * C::R - 1) / (C::THREADS * C::R);
  dim3 grid(nb, nchunks, 1);
  kern<<<grid, C::THREADS, C::SMEM, st>>>(a);
  return cudaGetLastError();
}

template <int CORE, int RES, bool FINE>
static int occupancy_one() {
  using C = PPConfig<CORE, RES, FINE>;
  auto kern = C::kernel();
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, C::SMEM) != cudaSuccess) return 1;
  return nb > 0 ? nb : 1;
}

template <int CORE, bool FINE>
static cudaError_t launch_pp_core_cfg(int res, const PPArgs& a, int nchunks, cudaStream_t st) {
  switch (res) {
    case RES_U:  return launch_one<CORE, RES_U, FINE>(a, nchunks, st);
    case RES_UW: return launch_one<CORE, RES_UW, FINE>(a, nchunks, st);
    case RES_UE: return launch_one<CORE, RES_UE, FINE>(a, nchunks, st);
    case RES_UG: return launch_one<CORE, RES_UG, FINE>(a, nchunks, st);
  }
  return cudaErrorInvalidValue;
}

template <int CORE>
cudaError_t launch_pp_core(int res, const PPArgs& a, int nchunks, cudaStream_t st, bool fine) {
  return fine ? launch_pp_core_cfg<CORE, true>(res, a, nchunks, st) : launch_pp_core_cfg<CORE, false>(res, a, nchunks, st);
}

template <int CORE, bool FINE>
static int pp_occupancy_core_cfg(int res) {
  switch (res) {
    case RES_U:  return occupancy_one<CORE, RES_U, FINE>();
    case RES_UW: return occupancy_one<CORE, RES_UW, FINE>();
    case RES_UE: return occupancy_one<CORE, RES_UE, FINE>();
    case RES_UG: return occupancy_one<CORE, RES_UG, FINE>();
  }
  return 1;
}

template <int CORE>
int pp_occupancy_core(int res, bool fine) {
  return fine ? pp_occupancy_core_cfg<CORE, true>(res) : pp_occupancy_core_cfg<CORE, false>(res);
}

}  // namespace o2d
