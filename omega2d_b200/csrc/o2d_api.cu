// o2d_api.cu -- the C ABI declared in include/omega2d_b200.h: context, device mirrors of
// host arrays, launch planning, NCCL source all-gather, measurement hooks.
#include "o2d_context.h"

#include <dlfcn.h>
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace o2d;

// ---- errors --------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int o2d_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---- NCCL through dlopen (no link-time dependency; single-GPU users never load it) -------------
struct NcclUid { char internal[128]; };
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (handle) break; }
    if (!handle) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
    AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && CommDestroy && AllGather && AllReduce && GetErrorString;
  }
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;
constexpr int kNcclChar = 0;     // ncclInt8 / ncclChar
constexpr int kNcclDouble = 8;   // ncclFloat64
constexpr int kNcclSum = 0;      // ncclSum

constexpr size_t kSmallCopy = 256 * 1024;        // arrays up to this size travel through the pinned buffer
constexpr size_t kPinBytes = 8 * 1024 * 1024;

#define bind_device o2d_bind_device
#define drain_timing o2d_drain_timing
int o2d_bind_device(o2d_context* c) {
  CU(cudaSetDevice(c->device));
  return O2D_OK;
}

int o2d_drain_timing(o2d_context* c) {
  for (auto& pr : c->ev_pending) {
    CU(cudaEventSynchronize(pr.second));
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, pr.first, pr.second));
    c->pair_ms += ms;
    c->ev_pool.push_back(pr);
  }
  c->ev_pending.clear();
  return O2D_OK;
}

PairTimer::PairTimer(o2d_context* ctx) : c(ctx) {
  if (!c->timing) return;
  if (c->ev_pending.size() >= 512) drain_timing(c);
  if (c->ev_pool.empty()) {
    cudaEventCreate(&pr.first);
    cudaEventCreate(&pr.second);
  } else {
    pr = c->ev_pool.back();
    c->ev_pool.pop_back();
  }
  cudaEventRecord(pr.first, c->stream);
}
PairTimer::~PairTimer() {
  if (!c->timing || !pr.first) return;
  cudaEventRecord(pr.second, c->stream);
  c->ev_pending.push_back(pr);
}

extern "C" const char* o2d_last_error(void) { return g_err; }

extern "C" int o2d_create(o2d_context** out, int device) {
  if (!out) return fail(O2D_ERR_INVALID, "o2d_create: ctx is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail(O2D_ERR_NO_DEVICE, "no CUDA device (%s); omega2d_b200 has no CPU fallback",
                e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(O2D_ERR_NO_DEVICE, "device %d out of range (have %d)", device, ndev);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(O2D_ERR_NO_DEVICE, "device %d is sm_%d%d; this library only carries sm_100a code", device,
                prop.major, prop.minor);
  o2d_context* c = new o2d_context();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  cudaDeviceGetAttribute(&c->sm_clock_khz, cudaDevAttrClockRate, device);
  CU(cudaSetDevice(device));
  CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  // pinned staging for small host arrays (see o2d_h2d_bytes); without it every copy takes the pageable path
  void* pin = nullptr;
  if (cudaHostAlloc(&pin, kPinBytes, cudaHostAllocDefault) == cudaSuccess) {
    c->pin = static_cast<char*>(pin);
    c->pin_cap = kPinBytes;
  } else {
    (void)cudaGetLastError();
  }
  *out = c;
  return O2D_OK;
}

extern "C" int o2d_comm_destroy(o2d_context* c);

extern "C" int o2d_destroy(o2d_context* c) {
  if (!c) return O2D_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  o2d_comm_destroy(c);
  for (DevBuf* b : {&c->packed, &c->packed_all, &c->partial, &c->panel_packed, &c->hsrc, &c->htgt, &c->hout, &c->hidx, &c->hpan, &c->adv, &c->coef, &c->counter, &c->vtu, &c->cil})
    b->release();
  if (c->vtu_host) cudaFreeHost(c->vtu_host);
  if (c->pin) cudaFreeHost(c->pin);
  for (auto& pr : c->ev_pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto& pr : c->ev_pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  cudaStreamDestroy(c->own_stream);
  delete c;
  return O2D_OK;
}

extern "C" int o2d_set_stream(o2d_context* c, void* s) {
  if (!c) return fail(O2D_ERR_INVALID, "ctx is NULL");
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return O2D_OK;
}

extern "C" int o2d_synchronize(o2d_context* c) {
  if (!c) return fail(O2D_ERR_INVALID, "ctx is NULL");
  if (int rc = bind_device(c)) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

// ---- launch planning -----------------------------------------------------------------------------------
// Pick the number of source chunks (grid.y) so that the last wave of CTAs is nearly full.
// nb target blocks x nchunks chunks run on `slots` = SMs x resident CTAs; a chunk costs
// ceil(ntiles/nchunks) tiles.  Efficiency = useful tile-slots / scheduled tile-slots.
static int choose_chunks(int nb, int ntiles, int slots, int max_chunks) {
  if (ntiles <= 1 || nb <= 0) return 1;
  int best = 1;
  double best_eff = 0.0;
  const int limit = std::min(ntiles, max_chunks);
  for (int nc = 1; nc <= limit; ++nc) {
    const int tpc = (ntiles + nc - 1) / nc;
    const int real_nc = (ntiles + tpc - 1) / tpc;
    if (real_nc != nc) continue;
    const long long items = (long long)nb * nc;
    const long long waves = (items + slots - 1) / slots;
    double eff = ((double)nb * ntiles) / ((double)waves * tpc * slots);
    // prefer fewer chunks at equal efficiency (each chunk re-reads targets and writes partials)
    eff -= 0.0005 * nc;
    if (eff > best_eff + 1e-9) { best_eff = eff; best = nc; }
    if (best_eff > 0.985) break;
  }
  return best;
}

int o2d_res_from_results(int results) {
  switch (results) {
    case O2D_VELONLY: return RES_U;
    case O2D_VELANDVORT: return RES_UW;
    case O2D_VELANDSHEAR: return RES_UE;
    case O2D_VELANDGRAD: return RES_UG;
  }
  return -1;
}

// ---- pair-sum launch plan ----------------------------------------------------------------------------------
// The (target block, source tile) steps of an evaluation, in block-major order, are cut into equal contiguous runs,
// one CTA each (PPArgs, o2d_internal.h).
//  * Kernels that run as ONE 512-thread CTA per SM get one run per SM.  Every SM has the same work, so there is no
//    last wave and no tail whatever the particle count, and only the two ends of a run touch the partial-sum buffer.
//  * Kernels with several CTAs per SM (the 24-warp shapes, and the "fine" configuration) get kRunsPerSlot short runs
//    per resident-CTA slot, placed by the hardware block scheduler.  Equal static runs do NOT work there: the warp
//    schedulers favour some resident CTAs over others (measured with three 256-thread CTAs per SM and one equal run
//    each: the three CTAs of an SM finished after 196, 240 and 323 ms; profiles/r02_pp_shape_search.md), whereas
//    the warps of ONE CTA are held together by the four-stage tile window.  The CTAs left on a draining SM speed up,
//    so the end-of-launch loss is ~1/10 of a run, not 1/2: 16 runs per slot cost ~0.5 %.
//  * "fine" configuration (128 threads, 128*R targets x 256 sources per step) when the coarse steps are too few to
//    give every SM a run worth starting (a few thousand to ~50 000 particles).
// O2D_PP_RUNS_PER_SLOT, O2D_PP_MIN_STEPS and O2D_PP_FINE_BELOW override the constants (tuning only).
struct PPPlan {
  bool fine;
  int ntiles, nblocks, nctas;
  long long steps_per_cta;
};
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}
static PPPlan pp_plan(o2d_context* c, int core, int res, long long ns_padded, int nt) {
  static const int kRunsPerSlot = std::max(1, env_int("O2D_PP_RUNS_PER_SLOT", 16));
  static const int kMinSteps = std::max(1, env_int("O2D_PP_MIN_STEPS", 2));
  static const int kFineBelow = std::max(0, env_int("O2D_PP_FINE_BELOW", 16));
  PPPlan p;
  const int btb = pp_block_targets(core, res, false);
  const long long big_steps = (long long)((nt + btb - 1) / btb) * (ns_padded / kBigTile);
  p.fine = big_steps < (long long)kFineBelow * c->sm_count;
  const int tile = p.fine ? kFineTile : kBigTile;
  const int bt = pp_block_targets(core, res, p.fine);
  p.ntiles = (int)(ns_padded / tile);
  p.nblocks = (nt + bt - 1) / bt;
  const long long total = (long long)p.nblocks * p.ntiles;
  const int occ = pp_occupancy(core, res, p.fine);
  const long long runs = (long long)c->sm_count * occ * (occ > 1 ? kRunsPerSlot : 1);
  p.steps_per_cta = std::max<long long>((total + runs - 1) / runs, occ > 1 ? std::min<long long>(kMinSteps, p.ntiles) : 1);
  p.nctas = (int)((total + p.steps_per_cta - 1) / p.steps_per_cta);
  return p;
}

// the pair sums of every tile of an already packed source array (ns_padded sources, a multiple of kTile) onto nt targets
static int pp_run_packed(o2d_context* c, int core, int res, const float4* packed, long long ns_padded,
                         const float* tx, const float* ty, const float* tr, int nt, const OutPtrs& out) {
  if (ns_padded <= 0 || nt <= 0) return O2D_OK;
  const PPPlan p = pp_plan(c, core, res, ns_padded, nt);
  PPArgs a;
  a.packed = packed;
  a.ntiles = p.ntiles;
  a.nblocks = p.nblocks;
  a.steps_per_cta = p.steps_per_cta;
  a.tx = tx; a.ty = ty; a.tr = tr; a.nt = nt;
  a.out = out;
  const bool cut = p.steps_per_cta % p.ntiles != 0;   // some blocks are shared between CTAs
  CU(c->partial.reserve((size_t)p.nctas * 2 * res_ncomp(res) * pp_block_targets(core, res, p.fine) * sizeof(double)));
  a.partial = c->partial.as<double>();
  {
    PairTimer timer(c);
    CU(launch_pp(core, res, a, p.nctas, c->stream, p.fine));
    c->launches++;
    c->pair_launches++;
  }
  if (cut) {
    CU(launch_pp_combine(core, res, a, c->stream, p.fine));
    c->launches++;
  }
  return O2D_OK;
}

int o2d_pp_device(o2d_context* c, int core, int res,
                  int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                  int nt, const float* tx, const float* ty, const float* tr, const OutPtrs& out) {
  if (ns <= 0 || nt <= 0) return O2D_OK;
  const int ns_padded = ((ns + kTile - 1) / kTile) * kTile;   // the packed array is always padded to whole big tiles
  CU(c->packed.reserve((size_t)ns_padded * 16));
  CU(launch_pack_sources(res == RES_VORT ? 100 : core, ns, sx, sy, ss, sr, c->packed.as<float4>(), ns_padded, c->stream));
  c->launches++;
  return pp_run_packed(c, core, res, c->packed.as<float4>(), ns_padded, tx, ty, tr, nt, out);
}

int o2d_pp_multi_device(o2d_context* c, int core, int res, int nsrc, const o2d_points_view* src,
                        int nt, const float* tx, const float* ty, const float* tr, const OutPtrs& out) {
  long long total = 0;
  int last = -1;
  for (int k = 0; k < nsrc; ++k) if (src[k].n > 0) { total += ((long long)src[k].n + 1) / 2 * 2; last = k; }
  if (last < 0 || nt <= 0) return O2D_OK;
  const long long ns_padded = (total + kTile - 1) / kTile * kTile;
  if (ns_padded > INT_MAX) return fail(O2D_ERR_INVALID, "more than 2^31 particles in one evaluation");
  CU(c->packed.reserve((size_t)ns_padded * 16));
  float4* packed = c->packed.as<float4>();
  const int mode = res == RES_VORT ? 100 : core;
  long long off = 0;
  for (int k = 0; k < nsrc; ++k) {
    if (src[k].n <= 0) continue;
    const int even = (src[k].n + 1) / 2 * 2;
    CU(launch_pack_sources(mode, src[k].n, src[k].x, src[k].y, src[k].s, src[k].r, packed + off, even, c->stream));
    c->launches++;
    off += even;
  }
  if (ns_padded > off) {   // tail of the last tile: zero-strength copies of the last collection's last particle
    CU(launch_pack_sources(mode, src[last].n, src[last].x, src[last].y, src[last].s, src[last].r, packed + off,
                           (int)(ns_padded - off), c->stream, src[last].n));
    c->launches++;
  }
  return pp_run_packed(c, core, res, packed, ns_padded, tx, ty, tr, nt, out);
}

// ---- host staging helpers ----------------------------------------------------------------------------------
// Host arrays are the caller's pageable std::vector storage.  A pageable cudaMemcpy is staged by the driver and
// costs ~10 us per call however small; an evaluation moves 10-15 arrays, which is most of the wall time of a
// 10k-particle call (C1/C2 sizes).  Arrays up to kSmallCopy bytes therefore go through the context's pinned
// buffer: memcpy + truly asynchronous DMA on the way in; DMA + memcpy after the final synchronise on the way out.
static char* pin_take(o2d_context* c, size_t bytes, size_t limit = kSmallCopy) {
  if (!c->pin || bytes > limit) return nullptr;
  const size_t need = (bytes + 63) / 64 * 64;
  if (c->pin_off + need > c->pin_cap) return nullptr;
  char* p = c->pin + c->pin_off;
  c->pin_off += need;
  return p;
}
// caller memory that is already page-locked (cudaHostAlloc / cudaHostRegister) needs no staging
static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
int o2d_h2d_bytes(o2d_context* c, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return O2D_OK;
  if (bytes <= kSmallCopy && is_pinned(src)) {
    // fall through to the direct asynchronous copy
  } else if (char* p = pin_take(c, bytes)) {
    memcpy(p, src, bytes);
    src = p;
  }
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return O2D_OK;
}
int o2d_d2h_bytes(o2d_context* c, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return O2D_OK;
  if (bytes <= kSmallCopy && is_pinned(dst)) {
    // direct
  } else if (char* p = pin_take(c, bytes)) {
    CU(cudaMemcpyAsync(p, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    c->pending_out.push_back({dst, p, bytes});
    return O2D_OK;
  }
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return O2D_OK;
}
int o2d_h2d(o2d_context* c, float* dst, const float* src, size_t n) { return o2d_h2d_bytes(c, dst, src, n * sizeof(float)); }
int o2d_d2h(o2d_context* c, float* dst, const float* src, size_t n) { return o2d_d2h_bytes(c, dst, src, n * sizeof(float)); }
int o2d_finish_host(o2d_context* c) {
  CU(cudaStreamSynchronize(c->stream));
  for (const auto& po : c->pending_out) memcpy(po.dst, po.src, po.bytes);
  c->pending_out.clear();
  c->pin_off = 0;
  return O2D_OK;
}
#define h2d o2d_h2d
#define d2h o2d_d2h

// span of a batch in device memory, if its items are ascending, non-overlapping and close together
static bool batch_span(const std::vector<HostBatch::Item>& items, char** base, size_t* span) {
  if (items.size() < 2) return false;
  char* lo = static_cast<char*>(items.front().dev);
  char* prev_end = lo;
  for (const auto& it : items) {
    char* d = static_cast<char*>(it.dev);
    if (d < prev_end || (size_t)(d - prev_end) > 1024) return false;   // only padding between neighbours
    prev_end = d + it.bytes;
  }
  *base = lo;
  *span = (size_t)(prev_end - lo);
  return true;
}
int HostBatch::upload() {
  char* base; size_t span;
  if (batch_span(items, &base, &span)) {
    if (char* p = pin_take(c, span, 4 * kSmallCopy)) {
      for (const auto& it : items) memcpy(p + (static_cast<char*>(it.dev) - base), it.host, it.bytes);
      CU(cudaMemcpyAsync(base, p, span, cudaMemcpyHostToDevice, c->stream));
      items.clear();
      return O2D_OK;
    }
  }
  for (const auto& it : items)
    if (int rc = o2d_h2d_bytes(c, it.dev, it.host, it.bytes)) return rc;
  items.clear();
  return O2D_OK;
}
int HostBatch::download() {
  char* base; size_t span;
  if (batch_span(items, &base, &span)) {
    if (char* p = pin_take(c, span, 4 * kSmallCopy)) {
      CU(cudaMemcpyAsync(p, base, span, cudaMemcpyDeviceToHost, c->stream));
      for (const auto& it : items) c->pending_out.push_back({it.host, p + (static_cast<char*>(it.dev) - base), it.bytes});
      items.clear();
      return O2D_OK;
    }
  }
  for (const auto& it : items)
    if (int rc = o2d_d2h_bytes(c, it.host, it.dev, it.bytes)) return rc;
  items.clear();
  return O2D_OK;
}

int o2d_check_common(o2d_context* c, int mem) {
  if (!c) return fail(O2D_ERR_INVALID, "ctx is NULL");
  if (mem != O2D_MEM_HOST && mem != O2D_MEM_DEVICE) return fail(O2D_ERR_INVALID, "unknown mem space %d", mem);
  if (mem == O2D_MEM_HOST) {       // a host-array call starts with an empty pinned buffer (an earlier call may have failed half way)
    if (!c->pending_out.empty()) { cudaStreamSynchronize(c->stream); c->pending_out.clear(); }
    c->pin_off = 0;
  }
  return bind_device(c);
}
#define check_common o2d_check_common

// gathers the output pointer list for a results type; returns number of outputs or <0
static int collect_outputs(int res, float* tu, float* tv, float* tq, float* tux, float* tvx, float* tuy,
                           float* tvy, float** outs) {
  if (!tu || !tv) return fail(O2D_ERR_INVALID, "tu/tv must not be NULL");
  outs[0] = tu; outs[1] = tv;
  if (res == RES_U) return 2;
  if (res == RES_UW || res == RES_UE) {
    if (!tq) return fail(O2D_ERR_INVALID, "tq must not be NULL for vorticity/shear results");
    outs[2] = tq;
    return 3;
  }
  if (!tux || !tvx || !tuy || !tvy) return fail(O2D_ERR_INVALID, "gradient outputs must not be NULL for velandgrad");
  outs[2] = tux; outs[3] = tvx; outs[4] = tuy; outs[5] = tvy;
  return 6;
}

extern "C" int o2d_points_on_points(o2d_context* c, int core, int results, int mem,
                                    int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                                    int nt, const float* tx, const float* ty, const float* tr,
                                    float* tu, float* tv, float* tq,
                                    float* tux, float* tvx, float* tuy, float* tvy) {
  if (int rc = check_common(c, mem)) return rc;
  if (core < 0 || core > 4) return fail(O2D_ERR_INVALID, "unknown core %d", core);
  if (results == O2D_PSIONLY)
    return fail(O2D_ERR_UNSUPPORTED, "Point elements cannot compute streamfunction yet (reference src/Influence.h:60)");
  const int res = o2d_res_from_results(results);
  if (res < 0) return fail(O2D_ERR_INVALID, "unknown results type %d", results);
  if (ns < 0 || nt < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (ns == 0 || nt == 0) return O2D_OK;
  if (!sx || !sy || !ss || !sr || !tx || !ty) return fail(O2D_ERR_INVALID, "NULL source/target array");
  float* outs[kMaxOut] = {nullptr};
  const int nout = collect_outputs(res, tu, tv, tq, tux, tvx, tuy, tvy, outs);
  if (nout < 0) return nout;
  OutPtrs o{};
  if (mem == O2D_MEM_DEVICE) {
    for (int k = 0; k < nout; ++k) o.p[k] = outs[k];
    return o2d_pp_device(c, core, res, ns, sx, sy, ss, sr, nt, tx, ty, tr, o);
  }
  // host arrays: mirror on the device, run, copy the accumulated outputs back
  CU(c->hsrc.reserve(4 * padded_floats(ns)));
  CU(c->htgt.reserve(3 * padded_floats(nt)));
  CU(c->hout.reserve(kMaxOut * padded_floats(nt)));
  Stager S(c, &c->hsrc), T(c, &c->htgt), O(c, &c->hout);
  float *dsx = S.take(ns), *dsy = S.take(ns), *dss = S.take(ns), *dsr = S.take(ns);
  float *dtx = T.take(nt), *dty = T.take(nt), *dtr = tr ? T.take(nt) : nullptr;
  int rc;
  HostBatch bs(c), bt(c), bo(c);
  bs.add(dsx, sx, ns); bs.add(dsy, sy, ns); bs.add(dss, ss, ns); bs.add(dsr, sr, ns);
  bt.add(dtx, tx, nt); bt.add(dty, ty, nt);
  if (tr) bt.add(dtr, tr, nt);
  for (int k = 0; k < nout; ++k) {
    o.p[k] = O.take(nt);
    bo.add(o.p[k], outs[k], nt);
  }
  if ((rc = bs.upload()) || (rc = bt.upload()) || (rc = bo.upload())) return rc;
  if ((rc = o2d_pp_device(c, core, res, ns, dsx, dsy, dss, dsr, nt, dtx, dty, dtr, o))) return rc;
  for (int k = 0; k < nout; ++k) bo.add(o.p[k], outs[k], nt);
  if ((rc = bo.download())) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

extern "C" int o2d_points_vorticity(o2d_context* c, int mem,
                                    int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                                    int nt, const float* tx, const float* ty, float* tw) {
  if (int rc = check_common(c, mem)) return rc;
  if (ns < 0 || nt < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (ns == 0 || nt == 0) return O2D_OK;
  if (!sx || !sy || !ss || !sr || !tx || !ty || !tw) return fail(O2D_ERR_INVALID, "NULL array");
  OutPtrs o{};
  if (mem == O2D_MEM_DEVICE) {
    o.p[0] = tw;
    return o2d_pp_device(c, CORE_RM, RES_VORT, ns, sx, sy, ss, sr, nt, tx, ty, nullptr, o);
  }
  CU(c->hsrc.reserve(4 * padded_floats(ns)));
  CU(c->htgt.reserve(3 * padded_floats(nt)));
  CU(c->hout.reserve(kMaxOut * padded_floats(nt)));
  Stager S(c, &c->hsrc), T(c, &c->htgt), O(c, &c->hout);
  float *dsx = S.take(ns), *dsy = S.take(ns), *dss = S.take(ns), *dsr = S.take(ns);
  float *dtx = T.take(nt), *dty = T.take(nt);
  o.p[0] = O.take(nt);
  int rc;
  if ((rc = h2d(c, dsx, sx, ns)) || (rc = h2d(c, dsy, sy, ns)) || (rc = h2d(c, dss, ss, ns)) ||
      (rc = h2d(c, dsr, sr, ns)) || (rc = h2d(c, dtx, tx, nt)) || (rc = h2d(c, dty, ty, nt)) ||
      (rc = h2d(c, o.p[0], tw, nt)))
    return rc;
  if ((rc = o2d_pp_device(c, CORE_RM, RES_VORT, ns, dsx, dsy, dss, dsr, nt, dtx, dty, nullptr, o))) return rc;
  if ((rc = d2h(c, tw, o.p[0], nt))) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

// ---- panels ------------------------------------------------------------------------------------------------
// panels already packed (np_padded of them, a multiple of kPanelTile) onto nt targets
static int panels_run_packed(o2d_context* c, int np_padded, int np_even, int nt, const float* tx, const float* ty, float* tu, float* tv) {
  const int ntiles = np_padded / kPanelTile;
  // two targets per thread halve the shared-memory reads per pair; one per thread when that would leave SMs idle
  const int rpt = (nt >= kPanBigR * 256 * c->sm_count * panels_occupancy(kPanBigR) || (nt >= 128 * kPanBigR * c->sm_count && ntiles >= 8)) ? kPanBigR : 1;
  const int nb = (nt + 128 * rpt - 1) / (128 * rpt);
  const int slots = c->sm_count * panels_occupancy(rpt);
  const int nchunks = choose_chunks(nb, ntiles, slots, 256);
  PanelArgs a;
  a.packed = c->panel_packed.as<float4>();
  a.npanels_padded = np_padded;
  a.npanels_even = np_even;
  a.tiles_per_chunk = (ntiles + nchunks - 1) / nchunks;
  a.tx = tx; a.ty = ty; a.nt = nt;
  a.out = OutPtrs{};
  a.out.p[0] = tu; a.out.p[1] = tv;
  a.partial = nullptr;
  if (nchunks > 1) {
    CU(c->partial.reserve((size_t)nchunks * 2 * (size_t)nt * sizeof(double)));
    a.partial = c->partial.as<double>();
  }
  {
    PairTimer timer(c);
    CU(launch_panels_on_points(a, nchunks, rpt, c->stream));
    c->launches++;
    c->pair_launches++;
  }
  if (nchunks > 1) {
    CU(launch_panels_combine(a.partial, nchunks, nt, a.out, c->stream));
    c->launches++;
  }
  return O2D_OK;
}

int o2d_panels_multi_device(o2d_context* c, int npan, const o2d_panels_view* pan, int nt, const float* tx, const float* ty,
                            float* tu, float* tv) {
  long long total = 0;
  int last = -1;
  for (int k = 0; k < npan; ++k) if (pan[k].np > 0) { total += ((long long)pan[k].np + 1) / 2 * 2; last = k; }
  if (last < 0 || nt <= 0) return O2D_OK;
  const long long np_padded = (total + kPanelTile - 1) / kPanelTile * kPanelTile;
  if (np_padded > INT_MAX) return fail(O2D_ERR_INVALID, "more than 2^31 panels in one evaluation");
  CU(c->panel_packed.reserve((size_t)np_padded * (16 * kPanelRecF4)));
  float4* packed = c->panel_packed.as<float4>();   // kPanelRecF4 float4 per panel
  long long off = 0;
  for (int k = 0; k < npan; ++k) {
    if (pan[k].np <= 0) continue;
    const int even = (pan[k].np + 1) / 2 * 2;
    CU(launch_pack_panels(pan[k].np, pan[k].nx, pan[k].ny, pan[k].idx, pan[k].vs, pan[k].ss, packed + kPanelRecF4 * off, even, c->stream));
    c->launches++;
    off += even;
  }
  if (np_padded > off) {
    CU(launch_pack_panels(pan[last].np, pan[last].nx, pan[last].ny, pan[last].idx, pan[last].vs, pan[last].ss, packed + kPanelRecF4 * off,
                          (int)(np_padded - off), c->stream, pan[last].np));
    c->launches++;
  }
  return panels_run_packed(c, (int)np_padded, (int)off, nt, tx, ty, tu, tv);
}

int o2d_panels_on_points_device(o2d_context* c, int np, const float* nx, const float* ny, const uint32_t* idx,
                                   const float* vs, const float* ss, int nt, const float* tx, const float* ty,
                                   float* tu, float* tv) {
  if (np <= 0 || nt <= 0) return O2D_OK;
  const int np_padded = (np + kPanelTile - 1) / kPanelTile * kPanelTile;
  CU(c->panel_packed.reserve((size_t)np_padded * (16 * kPanelRecF4)));
  CU(launch_pack_panels(np, nx, ny, idx, vs, ss, c->panel_packed.as<float4>(), np_padded, c->stream));
  c->launches++;
  return panels_run_packed(c, np_padded, (np + 1) / 2 * 2, nt, tx, ty, tu, tv);
}

extern "C" int o2d_panels_on_points(o2d_context* c, int mem,
                                    int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                    const float* vs, const float* ss,
                                    int nt, const float* tx, const float* ty, float* tu, float* tv) {
  if (int rc = check_common(c, mem)) return rc;
  if (np < 0 || nt < 0 || nnodes < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (np == 0 || nt == 0) return O2D_OK;
  if (!nx || !ny || !idx || !vs || !tx || !ty || !tu || !tv) return fail(O2D_ERR_INVALID, "NULL array");
  if (mem == O2D_MEM_DEVICE) return o2d_panels_on_points_device(c, np, nx, ny, idx, vs, ss, nt, tx, ty, tu, tv);
  for (int k = 0; k < 2 * np; ++k)
    if (idx[k] >= (uint32_t)nnodes) return fail(O2D_ERR_INVALID, "panel node index %u out of range", idx[k]);
  CU(c->hsrc.reserve(2 * padded_floats(nnodes) + 2 * padded_floats(np)));
  CU(c->hidx.reserve((size_t)2 * np * sizeof(uint32_t)));
  CU(c->htgt.reserve(3 * padded_floats(nt)));
  CU(c->hout.reserve(kMaxOut * padded_floats(nt)));
  Stager S(c, &c->hsrc), T(c, &c->htgt), O(c, &c->hout);
  float *dnx = S.take(nnodes), *dny = S.take(nnodes), *dvs = S.take(np), *dss = ss ? S.take(np) : nullptr;
  float *dtx = T.take(nt), *dty = T.take(nt);
  float *dtu = O.take(nt), *dtv = O.take(nt);
  int rc;
  if ((rc = h2d(c, dnx, nx, nnodes)) || (rc = h2d(c, dny, ny, nnodes)) || (rc = h2d(c, dvs, vs, np)) ||
      (rc = h2d(c, dtx, tx, nt)) || (rc = h2d(c, dty, ty, nt)) || (rc = h2d(c, dtu, tu, nt)) ||
      (rc = h2d(c, dtv, tv, nt)))
    return rc;
  if (ss && (rc = h2d(c, dss, ss, np))) return rc;
  if (int rc_ = o2d_h2d_bytes(c, c->hidx.p, idx, (size_t)2 * np * sizeof(uint32_t))) return rc_;
  if ((rc = o2d_panels_on_points_device(c, np, dnx, dny, c->hidx.as<uint32_t>(), dvs, dss, nt, dtx, dty, dtu, dtv)))
    return rc;
  if ((rc = d2h(c, tu, dtu, nt)) || (rc = d2h(c, tv, dtv, nt))) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

// `allreduce`: the particles are this rank's block of a sharded set -- the per-panel sums are reduced
// over all ranks (ncclAllReduce of 2*Np float64) before they are folded into pu,pv on every rank.
static int points_on_panels_impl(o2d_context* c, int ns, const float* sx, const float* sy, const float* ss,
                                 int np, const float* nx, const float* ny, const uint32_t* idx,
                                 const float* area, float* pu, float* pv, bool allreduce) {
  if (np <= 0) return O2D_OK;
  if (ns <= 0 && !allreduce) return O2D_OK;
  const int ntiles = (std::max(ns, 0) + kTile - 1) / kTile;
  const int ns_padded = ntiles * kTile;
  PtPanArgs a;
  a.ntiles = ntiles;
  a.nx = nx; a.ny = ny; a.idx = idx; a.area = area; a.np = np; a.pu = pu; a.pv = pv;
  a.partial = nullptr;
  a.packed_src = nullptr;
  int nchunks = 1;
  if (ntiles > 0) {
    CU(c->packed.reserve((size_t)ns_padded * 16));
    CU(launch_pack_sources(101, ns, sx, sy, ss, nullptr, c->packed.as<float4>(), ns_padded, c->stream));
    c->launches++;
    const int nb = (np + ptpan_block_panels() - 1) / ptpan_block_panels();
    const int slots = c->sm_count * ptpan_occupancy();
    nchunks = choose_chunks(nb, ntiles, slots, 4096);
    a.packed_src = c->packed.as<float4>();
  }
  a.tiles_per_chunk = ntiles > 0 ? (ntiles + nchunks - 1) / nchunks : 1;
  if (nchunks > 1 || allreduce) {
    // one extra [2][np] slot at the end receives the chunk-reduced (and rank-reduced) sums
    CU(c->partial.reserve(((size_t)nchunks + 1) * 2 * (size_t)np * sizeof(double)));
    a.partial = c->partial.as<double>();
  }
  if (ntiles > 0) {
    PairTimer timer(c);
    CU(launch_points_on_panels(a, nchunks, c->stream));
    c->launches++;
    c->pair_launches++;
  }
  if (allreduce) {
    double* red = a.partial + (size_t)nchunks * 2 * (size_t)np;
    CU(launch_ptpan_reduce_chunks(a.partial, ntiles > 0 ? nchunks : 0, np, red, c->stream));
    c->launches++;
    if (c->nranks > 1) {
      int r = g_nccl.AllReduce(red, red, (size_t)2 * np, kNcclDouble, kNcclSum, c->comm, c->stream);
      if (r != 0) return fail(O2D_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    }
    CU(launch_ptpan_combine_geom(red, 1, a, c->stream));
    c->launches++;
  } else if (nchunks > 1) {
    CU(launch_ptpan_combine_geom(a.partial, nchunks, a, c->stream));
    c->launches++;
  }
  return O2D_OK;
}

int o2d_points_on_panels_device(o2d_context* c, int ns, const float* sx, const float* sy, const float* ss,
                                int np, const float* nx, const float* ny, const uint32_t* idx,
                                const float* area, float* pu, float* pv) {
  return points_on_panels_impl(c, ns, sx, sy, ss, np, nx, ny, idx, area, pu, pv, false);
}

extern "C" int o2d_points_on_panels_sharded(o2d_context* c, int mem,
                                            int ns_local, const float* sx, const float* sy, const float* ss,
                                            int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                            const float* area, float* pu, float* pv) {
  if (int rc = check_common(c, mem)) return rc;
  if (c->nranks > 1 && !c->comm) return fail(O2D_ERR_COMM, "communicator not initialised (o2d_comm_init)");
  if (ns_local < 0 || np < 0 || nnodes < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (np == 0) return O2D_OK;
  if (ns_local > 0 && (!sx || !sy || !ss)) return fail(O2D_ERR_INVALID, "NULL particle array");
  if (!nx || !ny || !idx || !area || !pu || !pv) return fail(O2D_ERR_INVALID, "NULL array");
  if (mem == O2D_MEM_DEVICE) return points_on_panels_impl(c, ns_local, sx, sy, ss, np, nx, ny, idx, area, pu, pv, true);
  for (int k = 0; k < 2 * np; ++k)
    if (idx[k] >= (uint32_t)nnodes) return fail(O2D_ERR_INVALID, "panel node index %u out of range", idx[k]);
  CU(c->hsrc.reserve(3 * padded_floats(ns_local) + 256));
  CU(c->htgt.reserve(2 * padded_floats(nnodes) + padded_floats(np)));
  CU(c->hidx.reserve((size_t)2 * np * sizeof(uint32_t)));
  CU(c->hout.reserve(kMaxOut * padded_floats(np)));
  Stager S(c, &c->hsrc), T(c, &c->htgt), O(c, &c->hout);
  float *dsx = S.take(ns_local), *dsy = S.take(ns_local), *dss = S.take(ns_local);
  float *dnx = T.take(nnodes), *dny = T.take(nnodes), *darea = T.take(np);
  float *dpu = O.take(np), *dpv = O.take(np);
  int rc;
  if ((rc = h2d(c, dsx, sx, ns_local)) || (rc = h2d(c, dsy, sy, ns_local)) || (rc = h2d(c, dss, ss, ns_local)) ||
      (rc = h2d(c, dnx, nx, nnodes)) || (rc = h2d(c, dny, ny, nnodes)) || (rc = h2d(c, darea, area, np)) ||
      (rc = h2d(c, dpu, pu, np)) || (rc = h2d(c, dpv, pv, np)))
    return rc;
  if (int rc_ = o2d_h2d_bytes(c, c->hidx.p, idx, (size_t)2 * np * sizeof(uint32_t))) return rc_;
  if ((rc = points_on_panels_impl(c, ns_local, dsx, dsy, dss, np, dnx, dny, c->hidx.as<uint32_t>(), darea, dpu, dpv, true)))
    return rc;
  if ((rc = d2h(c, pu, dpu, np)) || (rc = d2h(c, pv, dpv, np))) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

extern "C" int o2d_points_on_panels(o2d_context* c, int mem,
                                    int ns, const float* sx, const float* sy, const float* ss,
                                    int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                    const float* area, float* pu, float* pv) {
  if (int rc = check_common(c, mem)) return rc;
  if (ns < 0 || np < 0 || nnodes < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (ns == 0 || np == 0) return O2D_OK;
  if (!sx || !sy || !ss || !nx || !ny || !idx || !area || !pu || !pv) return fail(O2D_ERR_INVALID, "NULL array");
  if (mem == O2D_MEM_DEVICE) return o2d_points_on_panels_device(c, ns, sx, sy, ss, np, nx, ny, idx, area, pu, pv);
  for (int k = 0; k < 2 * np; ++k)
    if (idx[k] >= (uint32_t)nnodes) return fail(O2D_ERR_INVALID, "panel node index %u out of range", idx[k]);
  CU(c->hsrc.reserve(3 * padded_floats(ns)));
  CU(c->htgt.reserve(2 * padded_floats(nnodes) + padded_floats(np)));
  CU(c->hidx.reserve((size_t)2 * np * sizeof(uint32_t)));
  CU(c->hout.reserve(kMaxOut * padded_floats(np)));
  Stager S(c, &c->hsrc), T(c, &c->htgt), O(c, &c->hout);
  float *dsx = S.take(ns), *dsy = S.take(ns), *dss = S.take(ns);
  float *dnx = T.take(nnodes), *dny = T.take(nnodes), *darea = T.take(np);
  float *dpu = O.take(np), *dpv = O.take(np);
  int rc;
  if ((rc = h2d(c, dsx, sx, ns)) || (rc = h2d(c, dsy, sy, ns)) || (rc = h2d(c, dss, ss, ns)) ||
      (rc = h2d(c, dnx, nx, nnodes)) || (rc = h2d(c, dny, ny, nnodes)) || (rc = h2d(c, darea, area, np)) ||
      (rc = h2d(c, dpu, pu, np)) || (rc = h2d(c, dpv, pv, np)))
    return rc;
  if (int rc_ = o2d_h2d_bytes(c, c->hidx.p, idx, (size_t)2 * np * sizeof(uint32_t))) return rc_;
  if ((rc = o2d_points_on_panels_device(c, ns, dsx, dsy, dss, np, dnx, dny, c->hidx.as<uint32_t>(), darea, dpu, dpv)))
    return rc;
  if ((rc = d2h(c, pu, dpu, np)) || (rc = d2h(c, pv, dpv, np))) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

// ---- zero / finalize -------------------------------------------------------------------------------------------
extern "C" int o2d_zero_vels(o2d_context* c, int mem, int n, float* u, float* v, float* w, float* e) {
  if (int rc = check_common(c, mem)) return rc;
  if (n < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (n == 0) return O2D_OK;
  if (!u || !v) return fail(O2D_ERR_INVALID, "NULL array");
  if (mem == O2D_MEM_HOST) {   // host arrays: nothing to launch, this is ElementBase::zero_vels' std::fill
    std::memset(u, 0, (size_t)n * sizeof(float));
    std::memset(v, 0, (size_t)n * sizeof(float));
    if (w) std::memset(w, 0, (size_t)n * sizeof(float));
    if (e) std::memset(e, 0, (size_t)n * sizeof(float));
    return O2D_OK;
  }
  CU(launch_zero(n, u, v, w, e, c->stream));
  c->launches++;
  return O2D_OK;
}

extern "C" int o2d_finalize_vels(o2d_context* c, int mem, int n, float* u, float* v, float* w, float* e,
                                 double fsx, double fsy) {
  if (int rc = check_common(c, mem)) return rc;
  if (n < 0) return fail(O2D_ERR_INVALID, "negative count");
  if (n == 0) return O2D_OK;
  if (!u || !v) return fail(O2D_ERR_INVALID, "NULL array");
  if (mem == O2D_MEM_DEVICE) {
    CU(launch_finalize(n, u, v, w, e, fsx, fsy, c->stream));
    c->launches++;
    return O2D_OK;
  }
  CU(c->hout.reserve(kMaxOut * padded_floats(n)));
  Stager O(c, &c->hout);
  float *du = O.take(n), *dv = O.take(n), *dw = w ? O.take(n) : nullptr, *de = e ? O.take(n) : nullptr;
  int rc;
  if ((rc = h2d(c, du, u, n)) || (rc = h2d(c, dv, v, n))) return rc;
  if (w && (rc = h2d(c, dw, w, n))) return rc;
  if (e && (rc = h2d(c, de, e, n))) return rc;
  CU(launch_finalize(n, du, dv, dw, de, fsx, fsy, c->stream));
  c->launches++;
  if ((rc = d2h(c, u, du, n)) || (rc = d2h(c, v, dv, n))) return rc;
  if (w && (rc = d2h(c, w, dw, n))) return rc;
  if (e && (rc = d2h(c, e, de, n))) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}

// ---- legacy plugin seam ------------------------------------------------------------------------------------------
static std::mutex g_legacy_mu;
static o2d_context* g_legacy_ctx = nullptr;
static o2d_group* g_legacy_group = nullptr;
static int g_default_core = -1;

static int default_core_locked() {
  if (g_default_core < 0) {
    g_default_core = O2D_CORE_V2;   // the reference's shipped #define (src/CoreFunc.h:38)
    if (const char* e = getenv("O2D_CORE")) {
      const char* names[] = {"rm", "exp", "wl", "v2", "v3"};
      for (int i = 0; i < 5; ++i)
        if (strcasecmp(e, names[i]) == 0) g_default_core = i;
    }
  }
  return g_default_core;
}
extern "C" int o2d_set_default_core(int core) {
  if (core < 0 || core > 4) return fail(O2D_ERR_INVALID, "unknown core %d", core);
  std::lock_guard<std::mutex> lk(g_legacy_mu);
  g_default_core = core;
  return O2D_OK;
}
extern "C" int o2d_get_default_core(void) {
  std::lock_guard<std::mutex> lk(g_legacy_mu);
  return default_core_locked();
}

static float legacy_solve(int* ns, const float* sx, const float* sy, const float* ss, const float* sr,
                          int* nt, const float* tx, const float* ty, const float* tr, float* tu, float* tv) {
  std::lock_guard<std::mutex> lk(g_legacy_mu);
  if (!ns || !nt) { fail(O2D_ERR_INVALID, "NULL count"); return -1.0f; }
  if (!g_legacy_ctx && !g_legacy_group) {
    int rc;
    if (const char* e = getenv("O2D_DEVICES")) {        // several GPUs from this one process (o2d_group.cu)
      rc = o2d_group_create(&g_legacy_group, strcasecmp(e, "all") == 0 ? 0 : atoi(e), nullptr);
    } else {
      int dev = 0;
      if (const char* e1 = getenv("O2D_DEVICE")) dev = atoi(e1);
      rc = o2d_create(&g_legacy_ctx, dev);
    }
    if (rc != O2D_OK) {
      fprintf(stderr, "omega2d_b200: %s\n", o2d_last_error());
      return -1.0f;
    }
  }
  const int core = default_core_locked();
  int rc = g_legacy_group
    ? o2d_group_points_on_points(g_legacy_group, core, O2D_VELONLY, *ns, sx, sy, ss, sr, *nt, tx, ty, tr,
                                 tu, tv, nullptr, nullptr, nullptr, nullptr, nullptr)
    : o2d_points_on_points(g_legacy_ctx, core, O2D_VELONLY, O2D_MEM_HOST, *ns, sx, sy, ss, sr, *nt, tx, ty, tr,
                           tu, tv, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc != O2D_OK) {
    fprintf(stderr, "omega2d_b200: %s\n", o2d_last_error());
    return -1.0f;
  }
  // flops as the reference's internal branch counts them (src/Influence.h:210,362)
  const float per_pair = (float)o2d_flops_per_pair(0, core, O2D_VELONLY, tr != nullptr);
  return (float)(*nt) * (2.0f + per_pair * (float)(*ns));
}

extern "C" float external_vel_solver_f_(int* ns, const float* sx, const float* sy, const float* ss, const float* sr,
                                        int* nt, const float* tx, const float* ty, float* tu, float* tv) {
  return legacy_solve(ns, sx, sy, ss, sr, nt, tx, ty, nullptr, tu, tv);
}
extern "C" float external_vel_solver_tr_f_(int* ns, const float* sx, const float* sy, const float* ss, const float* sr,
                                           int* nt, const float* tx, const float* ty, const float* tr,
                                           float* tu, float* tv) {
  return legacy_solve(ns, sx, sy, ss, sr, nt, tx, ty, tr, tu, tv);
}

// ---- flop accounting (reference counters) ---------------------------------------------------------------------------
extern "C" int o2d_flops_per_pair(int kind, int core, int results, int blob_targets) {
  // src/CoreFunc.h flops_tp_nograds / tv_nograds / tp_grads / tv_grads, per core RM EXP WL V2 V3
  static const int tp_nograds[5] = {3, 7, 5, 6, 8};
  static const int tv_nograds[5] = {5, 9, 7, 9, 12};
  static const int tp_grads[5] = {9, 14, 10, 10, 12};
  static const int tv_grads[5] = {9, 16, 12, 13, 16};
  if (kind == 1) return 28 + 3 + 8;   // flopsu_1v_0p, scalar flops_usf (src/Kernels.h:292,312)
  if (kind == 2) return 47;           // flopsu_1vs_0p (src/Kernels.h:362)
  if (kind == 3) return 13;           // src/InfluenceVort.h:102,119
  if (kind != 0 || core < 0 || core > 4) return -1;
  const int b = blob_targets ? 1 : 0;
  switch (results) {
    case O2D_VELONLY:     return 10 + (b ? tv_nograds[core] : tp_nograds[core]);   // src/Kernels.h:81,95
    case O2D_VELANDVORT:  return 19 + (b ? tv_grads[core] : tp_grads[core]);       // src/Kernels.h:139,160
    case O2D_VELANDSHEAR: return 23 + (b ? tv_grads[core] : tp_grads[core]);       // src/Kernels.h:185,209
    case O2D_VELANDGRAD:  return 25 + tv_grads[core];                              // src/Kernels.h:112
  }
  return -1;
}

// ---- measurement hooks -------------------------------------------------------------------------------------------------
extern "C" int o2d_enable_timing(o2d_context* c, int on) {
  if (!c) return fail(O2D_ERR_INVALID, "ctx is NULL");
  if (!on && c->timing) { if (int rc = drain_timing(c)) return rc; }
  c->timing = on != 0;
  return O2D_OK;
}
extern "C" int o2d_reset_counters(o2d_context* c) {
  if (!c) return fail(O2D_ERR_INVALID, "ctx is NULL");
  if (int rc = bind_device(c)) return rc;
  if (int rc = drain_timing(c)) return rc;
  c->launches = 0; c->pair_launches = 0; c->pair_ms = 0.0;
  return O2D_OK;
}
extern "C" int o2d_get_counters(o2d_context* c, long long* launches, long long* pair_launches, double* pair_ms) {
  if (!c) return fail(O2D_ERR_INVALID, "ctx is NULL");
  if (int rc = bind_device(c)) return rc;
  if (int rc = drain_timing(c)) return rc;
  if (launches) *launches = c->launches;
  if (pair_launches) *pair_launches = c->pair_launches;
  if (pair_ms) *pair_ms = c->pair_ms;
  return O2D_OK;
}
extern "C" int o2d_device_info(o2d_context* c, int* sm_count, int* sm_clock_mhz, double* fp32_peak_tflops) {
  if (!c) return fail(O2D_ERR_INVALID, "ctx is NULL");
  if (sm_count) *sm_count = c->sm_count;
  if (sm_clock_mhz) *sm_clock_mhz = c->sm_clock_khz / 1000;
  if (fp32_peak_tflops) *fp32_peak_tflops = (double)c->sm_count * 128.0 * 2.0 * (double)c->sm_clock_khz * 1e3 / 1e12;
  return O2D_OK;
}
extern "C" int o2d_measure_fp32_peak(o2d_context* c, int packed, double* tflops) {
  if (!c || !tflops) return fail(O2D_ERR_INVALID, "NULL argument");
  if (int rc = bind_device(c)) return rc;
  CU(c->partial.reserve(256));
  float* sink = c->partial.as<float>();
  const int blocks = c->sm_count * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(run_fp32_microbench(packed, blocks, threads, 64, sink, c->stream));   // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0, c->stream));
    CU(run_fp32_microbench(packed, blocks, threads, iters, sink, c->stream));
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * (double)kMicrobenchFmaPerIter * iters * (double)blocks * threads;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  c->launches += 6;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return O2D_OK;
}

// ---- multi-GPU ------------------------------------------------------------------------------------------------------------
extern "C" int o2d_comm_unique_id(void* id128) {
  if (!id128) return fail(O2D_ERR_INVALID, "id buffer is NULL");
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (!g_nccl.load()) return fail(O2D_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
  NcclUid uid;
  int r = g_nccl.GetUniqueId(&uid);
  if (r != 0) return fail(O2D_ERR_COMM, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
  static_assert(sizeof(NcclUid) == O2D_UNIQUE_ID_BYTES, "unique id size");
  memcpy(id128, &uid, sizeof(uid));
  return O2D_OK;
}

extern "C" int o2d_comm_init(o2d_context* c, int nranks, int rank, const void* id128) {
  if (!c || !id128) return fail(O2D_ERR_INVALID, "NULL argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(O2D_ERR_INVALID, "bad rank %d of %d", rank, nranks);
  if (int rc = bind_device(c)) return rc;
  {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (!g_nccl.load()) return fail(O2D_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
  }
  if (c->comm) o2d_comm_destroy(c);
  NcclUid uid;
  memcpy(&uid, id128, sizeof(uid));
  int r = g_nccl.CommInitRank(&c->comm, nranks, uid, rank);
  if (r != 0) { c->comm = nullptr; return fail(O2D_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
  c->nranks = nranks;
  c->rank = rank;
  return O2D_OK;
}

extern "C" int o2d_comm_destroy(o2d_context* c) {
  if (!c) return O2D_OK;
  if (c->comm) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    g_nccl.CommDestroy(c->comm);
    c->comm = nullptr;
  }
  c->nranks = 1;
  c->rank = 0;
  return O2D_OK;
}

// Multi-GPU: every rank packs its block of the particles into its slot of one gather buffer (blocks of ns_block
// entries, rounded up to a pair; shorter blocks are padded with zero-strength entries), ONE in-place ncclAllGather
// replicates the blocks (16 B per particle over NVLink), and the gathered array is then summed by the same single
// launch as on one GPU.  The gather sits on the launching stream in front of the sums: it moves 16 B per particle
// against >= 10^5 pair sums per particle, and a launch that owns the whole device with equal runs is worth more
// than the fraction of a millisecond an overlapped gather would hide.
int o2d_pp_sharded_device(o2d_context* c, int core, int res, int ns_local, int ns_block,
                             const float* sx, const float* sy, const float* ss, const float* sr,
                             int nt, const float* tx, const float* ty, const float* tr, const OutPtrs& out) {
  const int P = c->nranks;
  const long long blk = ((long long)ns_block + 1) / 2 * 2;                 // entries per rank block
  const long long ns_padded = (blk * P + kTile - 1) / kTile * kTile;       // gathered array, whole tiles
  if (ns_padded > INT_MAX) return fail(O2D_ERR_INVALID, "more than 2^31 particles in one evaluation");
  CU(c->packed_all.reserve((size_t)ns_padded * 16));
  float4* all = c->packed_all.as<float4>();
  float4* mine = all + (size_t)c->rank * blk;                              // 2 float4 per pair = 1 per entry
  CU(launch_pack_sources(core, ns_local, sx, sy, ss, sr, mine, (int)blk, c->stream));
  c->launches++;
  if (ns_padded > blk * P) {   // tail of the last tile: padding only
    CU(launch_pack_sources(core, ns_local, sx, sy, ss, sr, all + (size_t)blk * P, (int)(ns_padded - blk * P), c->stream, ns_local));
    c->launches++;
  }
  if (P > 1) {
    int r = g_nccl.AllGather(mine, all, (size_t)blk * 16, kNcclChar, c->comm, c->stream);
    if (r != 0) return fail(O2D_ERR_COMM, "ncclAllGather: %s", g_nccl.GetErrorString(r));
  }
  return pp_run_packed(c, core, res, all, ns_padded, tx, ty, tr, nt, out);
}

extern "C" int o2d_points_on_points_sharded(o2d_context* c, int core, int results, int mem,
                                            int ns_local, int ns_block,
                                            const float* sx, const float* sy, const float* ss, const float* sr,
                                            int nt, const float* tx, const float* ty, const float* tr,
                                            float* tu, float* tv, float* tq,
                                            float* tux, float* tvx, float* tuy, float* tvy) {
  if (int rc = check_common(c, mem)) return rc;
  if (c->nranks > 1 && !c->comm) return fail(O2D_ERR_COMM, "communicator not initialised (o2d_comm_init)");
  if (core < 0 || core > 4) return fail(O2D_ERR_INVALID, "unknown core %d", core);
  if (results == O2D_PSIONLY) return fail(O2D_ERR_UNSUPPORTED, "streamfunction results are not supported");
  const int res = o2d_res_from_results(results);
  if (res < 0) return fail(O2D_ERR_INVALID, "unknown results type %d", results);
  if (ns_local < 0 || nt < 0 || ns_block <= 0 || ns_local > ns_block)
    return fail(O2D_ERR_INVALID, "bad counts ns_local=%d ns_block=%d nt=%d", ns_local, ns_block, nt);
  if (ns_local > 0 && (!sx || !sy || !ss || !sr)) return fail(O2D_ERR_INVALID, "NULL source array");
  if (nt > 0 && (!tx || !ty)) return fail(O2D_ERR_INVALID, "NULL target array");
  float* outs[kMaxOut] = {nullptr};
  int nout = 0;
  if (nt > 0) {
    nout = collect_outputs(res, tu, tv, tq, tux, tvx, tuy, tvy, outs);
    if (nout < 0) return nout;
  }
  OutPtrs o{};
  if (mem == O2D_MEM_DEVICE) {
    for (int k = 0; k < nout; ++k) o.p[k] = outs[k];
    return o2d_pp_sharded_device(c, core, res, ns_local, ns_block, sx, sy, ss, sr, nt, tx, ty, tr, o);
  }
  CU(c->hsrc.reserve(4 * padded_floats(ns_local) + 256));
  CU(c->htgt.reserve(3 * padded_floats(nt) + 256));
  CU(c->hout.reserve(kMaxOut * padded_floats(nt) + 256));
  Stager S(c, &c->hsrc), T(c, &c->htgt), O(c, &c->hout);
  float *dsx = S.take(ns_local), *dsy = S.take(ns_local), *dss = S.take(ns_local), *dsr = S.take(ns_local);
  float *dtx = T.take(nt), *dty = T.take(nt), *dtr = tr ? T.take(nt) : nullptr;
  int rc;
  if ((rc = h2d(c, dsx, sx, ns_local)) || (rc = h2d(c, dsy, sy, ns_local)) || (rc = h2d(c, dss, ss, ns_local)) ||
      (rc = h2d(c, dsr, sr, ns_local)) || (rc = h2d(c, dtx, tx, nt)) || (rc = h2d(c, dty, ty, nt)))
    return rc;
  if (tr && (rc = h2d(c, dtr, tr, nt))) return rc;
  for (int k = 0; k < nout; ++k) {
    o.p[k] = O.take(nt);
    if ((rc = h2d(c, o.p[k], outs[k], nt))) return rc;
  }
  if ((rc = o2d_pp_sharded_device(c, core, res, ns_local, ns_block, dsx, dsy, dss, dsr, nt, dtx, dty, dtr, o))) return rc;
  for (int k = 0; k < nout; ++k)
    if ((rc = d2h(c, outs[k], o.p[k], nt))) return rc;
  if (int rc_ = o2d_finish_host(c)) return rc_;
  return O2D_OK;
}
