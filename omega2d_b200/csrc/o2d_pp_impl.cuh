// o2d_pp_impl.cuh -- the particles-on-points pair-sum kernel (one template, instantiated per core).
//
// Replaces the reference loop nests of points_affect_points (reference src/Influence.h:197-249
// inert targets, :347-403 blob targets) and points_affect_points_vorticity
// (src/InfluenceVort.h:106-118).
//
// Work decomposition
//   The launch's work is the sequence of (target block, source tile) steps in block-major order; every CTA
//   owns an equal contiguous run of it (PPArgs, o2d_internal.h), so all CTAs of a launch that fills the
//   machine once finish together whatever the particle count -- there is no last wave.  A thread owns R targets
//   in registers while its CTA walks the tiles of their block.  Source tiles (TILE sources = TILE*16 bytes)
//   are staged into shared memory by ONE TMA bulk copy per tile (cp.async.bulk + mbarrier), double buffered
//   across block boundaries too, and read back with warp-broadcast LDS.128.
// Arithmetic
//   Sources are stored pairwise -- {x0,x1,y0,y1} {s0,s1,rho0,rho1} -- so that one LDS.128 feeds
//   both lanes of Blackwell's packed FP32 instructions (FADD2/FMUL2/FFMA2): every FP32
//   instruction in the inner loop evaluates two source-target pairs.
// Accumulation (SURVEY 7 "hard parts": a single FP32 accumulator fails the 1e-5 tolerance)
//   per tile: float partial sums (two lanes x TILE/2 terms), folded into float64 once per tile;
//   the float64 sums of a block that lies inside one CTA's run are added straight into the float
//   outputs ("tu[i] += accum", as the reference); the pieces of a block cut between CTAs go to a
//   partials buffer that pp_combine_kernel sums in fixed CTA order -- results are bit-reproducible.
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
template <> struct ResTraits<RES_U>    { static constexpr int NCOMP = 2; static constexpr bool GRAD = false; };
template <> struct ResTraits<RES_UW>   { static constexpr int NCOMP = 4; static constexpr bool GRAD = true; };
template <> struct ResTraits<RES_UE>   { static constexpr int NCOMP = 3; static constexpr bool GRAD = true; };
template <> struct ResTraits<RES_UG>   { static constexpr int NCOMP = 6; static constexpr bool GRAD = true; };
template <> struct ResTraits<RES_VORT> { static constexpr int NCOMP = 1; static constexpr bool GRAD = false; };

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
    } else if constexpr (CORE == CORE_WL && UNI && !T::GRAD) {
      // Winckelmans-Leonard velocity with one radius term for all R targets: r2 = ss K = ss (r + s2 r^2) = r (ss + (ss s2) r)
      // with r = 1/D; ss*s2 is formed once per source pair, so a pair costs 8 lane-ops + 1 MUFU instead of 9
      const float2 ss2 = mul2(str, rho);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float2 dx = add2(tx[k], nsx);
        const float2 dy = add2(ty[k], nsy);
        const float2 r = rcp2(fma2(dy, dy, fma2(dx, dx, rho)));
        const float2 r2 = mul2(r, fma2(ss2, r, str));
        part[k][0] = fma2(r2, dy, part[k][0]);
        part[k][1] = fma2(r2, dx, part[k][1]);
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

#ifdef O2D_PP_DEBUG
// development only: per-CTA (smid, start, end) of the last pair-sum launch, read back by o2d_debug_cta_times
extern __device__ unsigned long long o2d_dbg_times[3 * 16384];
__device__ __forceinline__ unsigned long long dbg_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned dbg_smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#endif

// Pipeline: the CTA's warps share the staged tiles but never wait for each other.  Stage s is filled by a TMA bulk
// copy that completes on full[s]; every warp waits for it, sums the tile onto its own targets and then counts
// itself off on done[s]; the warp that counts off LAST refills the stage with the tile NS steps further along the
// run.  There is no CTA-wide barrier anywhere in the walk: a warp's targets, its block boundaries and its stores
// are its own business.  (Measured: the warps of an SM drift apart and stay apart, which is what keeps the FMA and
// the XU pipe busy at the same time -- warps that run in step all queue for the same pipe at the same moment; a
// version with one __syncthreads per tile and all CTAs starting together ran 12 % slower.)
// Resident warps per SM the register allocation must allow: 24 (<= 85 registers) for the velocity and Gaussian-
// vorticity kernels, 16 (<= 128) for the gradient-type results.  Left to itself ptxas takes 104-126 registers for
// some velocity instances and halves their occupancy.
// (shape per core and result kind: pp_targets / pp_min_warps / pp_unroll in o2d_internal.h.  Left to itself ptxas
// takes 104-126 registers for some instances and others spill at 80; the resident-warp bound pins the allocation.)
template <int CORE, int RES, int THREADS, int TILE, int R, int UNROLL>
__global__ void __launch_bounds__(THREADS, pp_min_warps(CORE, RES, THREADS == kFineThreads) * 32 / THREADS)
pp_kernel(const PPArgs a) {
  using T = ResTraits<RES>;
  constexpr int NC = T::NCOMP;
  constexpr int NS = kStages;
  constexpr int NWARPS = THREADS / 32;
  constexpr uint32_t TILE_BYTES = TILE * 16;
  constexpr int BT = THREADS * R;   // targets per block

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NS * TILE_BYTES);
  int* done = reinterpret_cast<int*>(smem_raw + NS * TILE_BYTES + NS * 8);
  auto stage = [&](int s) { return reinterpret_cast<const float4*>(smem_raw + (size_t)s * TILE_BYTES); };

  const int tid = threadIdx.x;
  const int lane = tid & 31;
#ifdef O2D_PP_DEBUG
  if (tid == 0 && blockIdx.x < 16384) { o2d_dbg_times[3 * blockIdx.x] = dbg_smid(); o2d_dbg_times[3 * blockIdx.x + 1] = dbg_globaltimer(); }
#endif
  // this CTA's run of (block, tile) steps
  const long long total = (long long)a.nblocks * a.ntiles;
  const long long run_begin = (long long)blockIdx.x * a.steps_per_cta;
  if (run_begin >= total) return;
  const int nsteps = (int)(min(run_begin + a.steps_per_cta, total) - run_begin);
  int block = (int)(run_begin / a.ntiles);
  int tile = (int)(run_begin - (long long)block * a.ntiles);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); done[s] = 0; }
    mbar_fence_init();
  }
  __syncthreads();   // the only CTA-wide barrier: the mbarriers exist before anyone polls them
  if (tid == 0) {    // fill the ring; the other warps are already fetching their targets
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (s < nsteps) {
        mbar_arrive_expect_tx(&full[s], TILE_BYTES);
        tma_load_1d(smem_raw + (size_t)s * TILE_BYTES, a.packed + (size_t)((tile + s) % a.ntiles) * TILE, TILE_BYTES, &full[s]);
      }
  }

  int step = 0;   // steps done by this warp; stage = step % NS, mbarrier phase = (step / NS) & 1
  for (int seg = 0; step < nsteps; ++seg, ++block, tile = 0) {
    const int seg_tiles = min(a.ntiles - tile, nsteps - step);

    // ---- this thread's targets ------------------------------------------------------------------
    const int tbase = block * BT + tid;
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
    // every lane's R targets share one radius term (inert targets; uniform-radius particle sets, the
    // reference's normal case): the warp takes the cheaper tile_sum<UNI>.  Out-of-range targets carry 0:
    // they only send the last warps of a blob-target launch down the general path.
    bool uni_t = true;
#pragma unroll
    for (int k = 1; k < R; ++k) uni_t = uni_t && (tau[k].x == tau[0].x);
    const bool uni = __all_sync(0xffffffffu, uni_t);

    double acc[R][NC];
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[k][c] = 0.0;

    // ---- walk the segment's source tiles ----------------------------------------------------------
    for (int j = 0; j < seg_tiles; ++j, ++step) {
      const int st = step % NS;
      mbar_wait(&full[st], (uint32_t)((step / NS) & 1));
      const float4* s = stage(st);

      // EXP core: largest source radius term of the tile (radius terms are >= 0)
      float rho_max = 0.0f;
      if (CORE == CORE_EXP && RES != RES_VORT) {
        for (int p = lane; p < TILE / 2; p += 32) { const float4 B = s[2 * p + 1]; rho_max = fmaxf(rho_max, fmaxf(B.z, B.w)); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) rho_max = fmaxf(rho_max, __shfl_xor_sync(0xffffffffu, rho_max, off));
      }

      float2 part[R][NC];
#pragma unroll
      for (int k = 0; k < R; ++k)
#pragma unroll
        for (int c = 0; c < NC; ++c) part[k][c] = f2(0.0f);

      if (uni) tile_sum<CORE, RES, TILE, R, UNROLL, true>(s, tx, ty, tau, part, rho_max);
      else     tile_sum<CORE, RES, TILE, R, UNROLL, false>(s, tx, ty, tau, part, rho_max);

      // this warp is done with the stage; the last warp to say so refills it
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(&done[st], 1) == NWARPS - 1) {
          done[st] = 0;
          if (step + NS < nsteps) {
            mbar_arrive_expect_tx(&full[st], TILE_BYTES);
            tma_load_1d(smem_raw + (size_t)st * TILE_BYTES, a.packed + (size_t)((tile + j + NS) % a.ntiles) * TILE,
                        TILE_BYTES, &full[st]);
          }
        }
      }
      // fold this tile's float partial sums into the float64 accumulators
#pragma unroll
      for (int k = 0; k < R; ++k)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const double sum = (double)part[k][c].x + (double)part[k][c].y;
          acc[k][c] += bb_fold_scale<CORE, RES>(c) * sum;
        }
    }

    // ---- end of the segment -------------------------------------------------------------------------
#ifdef O2D_PP_DEBUG
    if (tid == THREADS - 1 && blockIdx.x < 16384) o2d_dbg_times[3 * blockIdx.x + 2] = dbg_globaltimer();
#endif
    if (seg_tiles == a.ntiles) {
      // the whole block was summed here: "tu[i] += accum" as the reference does
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int i = tbase + k * THREADS;
        if (i < a.nt) apply_outputs<RES>(a.out, i, acc[k]);
      }
    } else {
      // a run begins and ends with a piece of a block at most: slot 0 = the piece in the run's first block, slot 1 = the
      // other.  partial[((cta * 2 + slot) * NC + c) * BT + k*THREADS + tid], coalesced over tid
      double* dst = a.partial + ((size_t)blockIdx.x * 2 + (seg == 0 ? 0 : 1)) * (size_t)(NC * BT) + tid;
#pragma unroll
      for (int k = 0; k < R; ++k)
#pragma unroll
        for (int c = 0; c < NC; ++c) dst[c * BT + k * THREADS] = acc[k][c];
    }
  }
}

// Blocks that were cut between CTAs: sum the pieces in CTA order (fixed by the launch geometry, not by scheduling),
// then fold into the outputs.  Block b = steps [b*ntiles, (b+1)*ntiles) lies in the runs of CTAs
// k0 = b*ntiles/S .. k1 = ((b+1)*ntiles-1)/S; k0 == k1 means one CTA had all of it and has already added it.
template <int RES>
__global__ void __launch_bounds__(256)
pp_combine_kernel(const PPArgs a, int BT) {
  constexpr int NC = ResTraits<RES>::NCOMP;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nt) return;
  const int b = i / BT, li = i - b * BT;
  const long long S = a.steps_per_cta;
  const long long first = (long long)b * a.ntiles;
  const int k0 = (int)(first / S), k1 = (int)((first + a.ntiles - 1) / S);
  if (k0 == k1) return;
  double c[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) c[k] = 0.0;
  for (int k = k0; k <= k1; ++k) {
    const int slot = (int)(((long long)k * S) / a.ntiles) == b ? 0 : 1;
    const double* src = a.partial + ((size_t)k * 2 + slot) * (size_t)(NC * BT) + li;
#pragma unroll
    for (int q = 0; q < NC; ++q) c[q] += src[q * BT];
  }
  apply_outputs<RES>(a.out, i, c);
}

}  // namespace o2d

// ---- per-core launcher (explicitly instantiated in o2d_pp_<core>.cu) ---------------------------
namespace o2d {

template <int CORE, int RES, bool FINE>
struct PPConfig {
  static constexpr int THREADS = pp_threads(CORE, RES, FINE);
  static constexpr int TILE = FINE ? kFineTile : kBigTile;
  static constexpr int R = pp_targets(CORE, RES, FINE);
  static constexpr size_t SMEM = kStages * ((size_t)TILE * 16 + 8 + 4);
  static constexpr int UNROLL = pp_unroll(CORE, RES, FINE);
  static auto kernel() { return pp_kernel<CORE, RES, THREADS, TILE, R, UNROLL>; }
};

template <int CORE, int RES, bool FINE>
static cudaError_t launch_one(const PPArgs& a, int nctas, cudaStream_t st) {
  using C = PPConfig<CORE, RES, FINE>;
  auto kern = C::kernel();
  // four 8 KiB stages + barriers: under the 48 KiB a kernel gets without opting in, so there is no per-device
  // function attribute to set (a process may drive several devices, o2d_group.cu)
  static_assert(C::SMEM <= 48 * 1024, "dynamic shared memory above the default limit needs cudaFuncSetAttribute per device");
  kern<<<nctas, C::THREADS, C::SMEM, st>>>(a);
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
static cudaError_t launch_pp_core_cfg(int res, const PPArgs& a, int nctas, cudaStream_t st) {
  switch (res) {
    case RES_U:  return launch_one<CORE, RES_U, FINE>(a, nctas, st);
    case RES_UW: return launch_one<CORE, RES_UW, FINE>(a, nctas, st);
    case RES_UE: return launch_one<CORE, RES_UE, FINE>(a, nctas, st);
    case RES_UG: return launch_one<CORE, RES_UG, FINE>(a, nctas, st);
  }
  return cudaErrorInvalidValue;
}

template <int CORE>
cudaError_t launch_pp_core(int res, const PPArgs& a, int nctas, cudaStream_t st, bool fine) {
  return fine ? launch_pp_core_cfg<CORE, true>(res, a, nctas, st) : launch_pp_core_cfg<CORE, false>(res, a, nctas, st);
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
