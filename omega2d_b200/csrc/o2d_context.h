// o2d_context.h -- private: the context object and host-side helpers shared by the translation units
// that implement the C ABI (o2d_api.cu, o2d_substep.cu, o2d_bem.cu, o2d_reflect.cu, o2d_status.cu, o2d_vtu.cu).  Not installed.
#pragma once
#include "../../include/omega2d_b200.h"
#include "o2d_internal.h"

#include <cstdarg>
#include <cstdio>
#include <utility>
#include <vector>

// ---- errors --------------------------------------------------------------------------------------
int o2d_fail(int code, const char* fmt, ...);
#define fail o2d_fail
#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(O2D_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ---- grow-only device buffer -----------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaError_t e = cudaFree(p); if (e != cudaSuccess) return e; p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return e;
    cap = want;
    return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

typedef struct ncclComm* ncclComm_t;

// ---- context -----------------------------------------------------------------------------------------
struct o2d_context {
  int device = 0;
  int sm_count = 0;
  int sm_clock_khz = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;   // stream in use (own or caller's)
  // scratch
  DevBuf packed, packed_all, partial, panel_packed;
  DevBuf hsrc, htgt, hout, hidx, hpan;   // device mirrors for O2D_MEM_HOST calls
  DevBuf adv;                            // o2d_advect: intermediate Runge-Kutta stages
  DevBuf counter;                        // small device-side counters
  DevBuf coef;                           // o2d_panels_on_panels_coeff: device copy of the matrix block
  DevBuf vtu;                            // o2d_vtu_points: base64 staging
  DevBuf cil;                            // clear_inner_layer / reflect_interior: node normals, packed panels, group bounds
  int cil_np = 0;
  size_t cil_nn_bytes = 0;
  void* vtu_host = nullptr;              // o2d_write_vtu_points: pinned file image
  size_t vtu_host_cap = 0;
  // small host arrays travel through this pinned buffer (see o2d_h2d_bytes): a pageable cudaMemcpy costs
  // ~10 us of driver staging per call whatever its size, which dominates 10k-particle evaluations
  char* pin = nullptr;
  size_t pin_cap = 0, pin_off = 0;
  struct PendingOut { void* dst; const void* src; size_t bytes; };
  std::vector<PendingOut> pending_out;   // D2H copies parked in `pin`, delivered by o2d_finish_host
  // comm
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  // counters
  long long launches = 0, pair_launches = 0;
  double pair_ms = 0.0;
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool, ev_pending;
};

int o2d_bind_device(o2d_context* c);
int o2d_drain_timing(o2d_context* c);

struct PairTimer {   // brackets the dominant kernel launches with events on the launching stream
  o2d_context* c;
  std::pair<cudaEvent_t, cudaEvent_t> pr{nullptr, nullptr};
  explicit PairTimer(o2d_context* ctx);
  ~PairTimer();
};

// ---- host staging helpers ----------------------------------------------------------------------------------
struct Stager {   // carve float arrays out of one device buffer
  o2d_context* c;
  DevBuf* buf;
  size_t off = 0;
  Stager(o2d_context* ctx, DevBuf* b) : c(ctx), buf(b) {}
  float* take(size_t n) {
    float* p = reinterpret_cast<float*>(reinterpret_cast<char*>(buf->p) + off);
    off += ((n * sizeof(float) + 255) / 256) * 256;
    return p;
  }
};
inline size_t padded_floats(size_t n) { return ((n * sizeof(float) + 255) / 256) * 256; }

// Several small host arrays that land next to each other in one device buffer (Stager layout) travel as ONE DMA:
// each is memcpy'd to the same relative offset in the context's pinned buffer, then one cudaMemcpyAsync moves the
// span.  A 10 000-particle evaluation moves 7-11 arrays; at ~4 us of driver time per copy that is a third of the call.
// Falls back to one copy per array when the span does not fit the pinned buffer (large arrays: bandwidth-bound anyway).
struct HostBatch {
  o2d_context* c;
  struct Item { void* dev; void* host; size_t bytes; };
  std::vector<Item> items;
  explicit HostBatch(o2d_context* ctx) : c(ctx) {}
  void add(void* dev, const void* host, size_t bytes) { if (bytes) items.push_back({dev, const_cast<void*>(host), bytes}); }
  void add(float* dev, const float* host, size_t n) { add((void*)dev, (const void*)host, n * sizeof(float)); }
  int upload();     // host -> device, asynchronous on c->stream
  int download();   // device -> host; the data is in place after o2d_finish_host
};
int o2d_h2d_bytes(o2d_context* c, void* dst, const void* src, size_t bytes);
int o2d_d2h_bytes(o2d_context* c, void* dst, const void* src, size_t bytes);
int o2d_h2d(o2d_context* c, float* dst, const float* src, size_t n);
int o2d_d2h(o2d_context* c, float* dst, const float* src, size_t n);
// end of a host-array call: wait for the stream, then hand over the small results parked in the pinned buffer
int o2d_finish_host(o2d_context* c);
int o2d_check_common(o2d_context* c, int mem);

// ---- device-pointer workers (all arrays in device memory, asynchronous on c->stream) --------------------
int o2d_res_from_results(int results);
int o2d_pp_device(o2d_context* c, int core, int res,
                  int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                  int nt, const float* tx, const float* ty, const float* tr, const o2d::OutPtrs& out);
int o2d_pp_sharded_device(o2d_context* c, int core, int res, int ns_local, int ns_block,
                          const float* sx, const float* sy, const float* ss, const float* sr,
                          int nt, const float* tx, const float* ty, const float* tr, const o2d::OutPtrs& out);
int o2d_panels_on_points_device(o2d_context* c, int np, const float* nx, const float* ny, const uint32_t* idx,
                                const float* vs, const float* ss, int nt, const float* tx, const float* ty,
                                float* tu, float* tv);
int o2d_points_on_panels_device(o2d_context* c, int ns, const float* sx, const float* sy, const float* ss,
                                int np, const float* nx, const float* ny, const uint32_t* idx,
                                const float* area, float* pu, float* pv);
// clear_inner_layer push (mode 1) / reflect_interior (mode 2) on device arrays; d_moved may be nullptr
int o2d_cil_prepare(o2d_context* c, int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                    const float* snx, const float* sny);
int o2d_cil_run(o2d_context* c, int mode, int n, float* tx, float* ty, float cutoff_mult, float ips, unsigned long long* d_moved);
int o2d_cil_device(o2d_context* c, int mode, int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                   const float* snx, const float* sny, int n, float* tx, float* ty, float cutoff_mult, float ips,
                   unsigned long long* d_moved);
// several source collections behind one another: ONE pair-sum launch / ONE panel launch (o2d_find_vels_multi)
int o2d_pp_multi_device(o2d_context* c, int core, int res, int nsrc, const o2d_points_view* src,
                        int nt, const float* tx, const float* ty, const float* tr, const o2d::OutPtrs& out);
int o2d_panels_multi_device(o2d_context* c, int npan, const o2d_panels_view* pan, int nt, const float* tx, const float* ty,
                            float* tu, float* tv);
