// o2d_group.cu -- all GPUs of one box from ONE host process (SURVEY 8e "process model").
//
// Omega2D itself is a single-process OpenMP program: its influence loops are "#pragma omp parallel for"
// over the targets (reference src/Influence.h:198,348).  A group is the same decomposition over devices:
// the target collection is cut into one contiguous block per GPU, every GPU receives the whole source
// collection from host memory over its own PCIe link, and one host thread per device drives the ordinary
// single-device entry point on its block, so every result is produced by exactly the kernels (and the
// launch plan rules) of the single-GPU path.  No collective is needed: sources come from the host, outputs
// go back to disjoint slices of the caller's arrays.  (The one-process-per-GPU variant with device-resident
// particles and an NCCL all-gather is o2d_points_on_points_sharded, o2d_api.cu.)
// points_affect_panels is the opposite case -- few targets, many sources -- so there the SOURCES are cut
// into blocks and the per-panel results of the devices are added on the host.
#include "o2d_context.h"

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

struct o2d_group {
  std::vector<o2d_context*> ctx;
  long long min_pairs_per_device = 2000000000ll;   // below this a block is not worth a second device
};

namespace {

struct Block { int lo, n; };

// contiguous blocks, multiples of 1024 elements (one CTA of the pair kernel covers 1024 targets)
std::vector<Block> cut(int n, int parts) {
  std::vector<Block> b;
  if (n <= 0) return b;
  parts = std::max(1, parts);
  int per = (n + parts - 1) / parts;
  per = (per + 1023) / 1024 * 1024;
  for (int lo = 0; lo < n; lo += per) b.push_back({lo, std::min(per, n - lo)});
  return b;
}

int devices_for(const o2d_group* g, double pairs) {
  const long long k = (long long)(pairs / (double)g->min_pairs_per_device);
  return (int)std::max(1ll, std::min<long long>((long long)g->ctx.size(), k));
}

// run fn(device slot, block) on one host thread per block; first failure wins
template <class F>
int run_blocks(const std::vector<Block>& blocks, F fn) {
  const int nb = (int)blocks.size();
  if (nb == 0) return O2D_OK;
  std::vector<int> rc(nb, O2D_OK);
  std::vector<std::string> msg(nb);
  auto body = [&](int k) {
    rc[k] = fn(k, blocks[k]);
    if (rc[k] != O2D_OK) msg[k] = o2d_last_error();   // the message is thread-local: carry it over
  };
  std::vector<std::thread> th;
  int started = 1;
  try {
    for (int k = 1; k < nb; ++k) { th.emplace_back(body, k); ++started; }
  } catch (...) {
    // no more host threads to be had: the remaining blocks run on this thread, nothing crosses the C ABI
  }
  body(0);
  for (int k = started; k < nb; ++k) body(k);
  for (auto& t : th) t.join();
  for (int k = 0; k < nb; ++k)
    if (rc[k] != O2D_OK) return fail(rc[k], "device slot %d: %s", k, msg[k].c_str());
  return O2D_OK;
}

inline const float* at(const float* p, int lo) { return p ? p + lo : nullptr; }
inline float* at(float* p, int lo) { return p ? p + lo : nullptr; }

}  // namespace

extern "C" int o2d_group_create(o2d_group** out, int ndev, const int* devices) {
  if (!out) return fail(O2D_ERR_INVALID, "o2d_group_create: group is NULL");
  *out = nullptr;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0)
    return fail(O2D_ERR_NO_DEVICE, "no CUDA device visible: omega2d_b200 has no CPU fallback");
  if (ndev <= 0) { ndev = visible; devices = nullptr; }
  o2d_group* g = new o2d_group;
  for (int k = 0; k < ndev; ++k) {
    o2d_context* c = nullptr;
    const int rc = o2d_create(&c, devices ? devices[k] : k);
    if (rc != O2D_OK) {
      for (o2d_context* p : g->ctx) o2d_destroy(p);
      delete g;
      return rc;
    }
    g->ctx.push_back(c);
  }
  *out = g;
  return O2D_OK;
}

extern "C" int o2d_group_destroy(o2d_group* g) {
  if (!g) return O2D_OK;
  for (o2d_context* c : g->ctx) o2d_destroy(c);
  delete g;
  return O2D_OK;
}

extern "C" int o2d_group_size(const o2d_group* g) { return g ? (int)g->ctx.size() : 0; }

extern "C" o2d_context* o2d_group_context(o2d_group* g, int slot) {
  if (!g || slot < 0 || slot >= (int)g->ctx.size()) return nullptr;
  return g->ctx[slot];
}

extern "C" int o2d_group_set_min_pairs(o2d_group* g, long long pairs_per_device) {
  if (!g || pairs_per_device < 0) return fail(O2D_ERR_INVALID, "bad argument");
  g->min_pairs_per_device = std::max(1ll, pairs_per_device);
  return O2D_OK;
}

extern "C" int o2d_group_points_on_points(o2d_group* g, int core, int results,
                                          int ns, const float* sx, const float* sy, const float* ss, const float* sr,
                                          int nt, const float* tx, const float* ty, const float* tr,
                                          float* tu, float* tv, float* tq,
                                          float* tux, float* tvx, float* tuy, float* tvy) {
  if (!g || g->ctx.empty()) return fail(O2D_ERR_INVALID, "group is NULL");
  const auto blocks = cut(nt, devices_for(g, (double)ns * (double)nt));
  if (blocks.size() <= 1)   // also the path that validates arguments for empty inputs
    return o2d_points_on_points(g->ctx[0], core, results, O2D_MEM_HOST, ns, sx, sy, ss, sr, nt, tx, ty, tr,
                                tu, tv, tq, tux, tvx, tuy, tvy);
  if (!tx || !ty || !tu || !tv) return fail(O2D_ERR_INVALID, "NULL target array");   // offsets below would hide a NULL
  return run_blocks(blocks, [&](int k, Block b) {
    return o2d_points_on_points(g->ctx[k], core, results, O2D_MEM_HOST, ns, sx, sy, ss, sr,
                                b.n, tx + b.lo, ty + b.lo, at(tr, b.lo), at(tu, b.lo), at(tv, b.lo), at(tq, b.lo),
                                at(tux, b.lo), at(tvx, b.lo), at(tuy, b.lo), at(tvy, b.lo));
  });
}

extern "C" int o2d_group_find_vels(o2d_group* g, int core, int results,
                                   int nv, const float* vx, const float* vy, const float* vs, const float* vr,
                                   int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                   const float* pvs, const float* pss,
                                   int nt, const float* tx, const float* ty, const float* tr,
                                   float* tu, float* tv, float* tq, double fsx, double fsy) {
  if (!g || g->ctx.empty()) return fail(O2D_ERR_INVALID, "group is NULL");
  // a panel pair costs about five particle pairs
  const auto blocks = cut(nt, devices_for(g, ((double)nv + 5.0 * (double)np) * (double)nt));
  if (blocks.size() <= 1)
    return o2d_find_vels(g->ctx[0], core, results, O2D_MEM_HOST, nv, vx, vy, vs, vr, np, nnodes, nx, ny, idx, pvs, pss,
                         nt, tx, ty, tr, tu, tv, tq, fsx, fsy);
  if (!tx || !ty || !tu || !tv) return fail(O2D_ERR_INVALID, "NULL target array");
  return run_blocks(blocks, [&](int k, Block b) {
    return o2d_find_vels(g->ctx[k], core, results, O2D_MEM_HOST, nv, vx, vy, vs, vr, np, nnodes, nx, ny, idx, pvs, pss,
                         b.n, tx + b.lo, ty + b.lo, at(tr, b.lo), at(tu, b.lo), at(tv, b.lo), at(tq, b.lo), fsx, fsy);
  });
}

extern "C" int o2d_group_panels_on_points(o2d_group* g, int np, int nnodes, const float* nx, const float* ny,
                                          const uint32_t* idx, const float* vs, const float* ss,
                                          int nt, const float* tx, const float* ty, float* tu, float* tv) {
  if (!g || g->ctx.empty()) return fail(O2D_ERR_INVALID, "group is NULL");
  const auto blocks = cut(nt, devices_for(g, 5.0 * (double)np * (double)nt));
  if (blocks.size() <= 1)
    return o2d_panels_on_points(g->ctx[0], O2D_MEM_HOST, np, nnodes, nx, ny, idx, vs, ss, nt, tx, ty, tu, tv);
  if (!tx || !ty || !tu || !tv) return fail(O2D_ERR_INVALID, "NULL target array");
  return run_blocks(blocks, [&](int k, Block b) {
    return o2d_panels_on_points(g->ctx[k], O2D_MEM_HOST, np, nnodes, nx, ny, idx, vs, ss,
                                b.n, tx + b.lo, ty + b.lo, tu + b.lo, tv + b.lo);
  });
}

extern "C" int o2d_group_points_on_panels(o2d_group* g, int ns, const float* sx, const float* sy, const float* ss,
                                          int np, int nnodes, const float* nx, const float* ny, const uint32_t* idx,
                                          const float* area, float* pu, float* pv) {
  if (!g || g->ctx.empty()) return fail(O2D_ERR_INVALID, "group is NULL");
  // sources are cut; every device computes "0 - sum_block/area" for all panels, the host adds the slices up
  const auto blocks = cut(ns, devices_for(g, 5.0 * (double)np * (double)ns));
  if (blocks.size() <= 1 || np <= 0)
    return o2d_points_on_panels(g->ctx[0], O2D_MEM_HOST, ns, sx, sy, ss, np, nnodes, nx, ny, idx, area, pu, pv);
  if (!pu || !pv || !sx || !sy || !ss) return fail(O2D_ERR_INVALID, "NULL source or output array");
  std::vector<float> part(2 * (size_t)np * blocks.size(), 0.0f);
  const int rc = run_blocks(blocks, [&](int k, Block b) {
    float* u = part.data() + 2 * (size_t)np * k;
    return o2d_points_on_panels(g->ctx[k], O2D_MEM_HOST, b.n, sx + b.lo, sy + b.lo, ss + b.lo, np, nnodes, nx, ny, idx,
                                area, u, u + np);
  });
  if (rc != O2D_OK) return rc;
  for (int i = 0; i < np; ++i) {
    double du = 0.0, dv = 0.0;   // fixed block order
    for (size_t k = 0; k < blocks.size(); ++k) {
      du += (double)part[2 * (size_t)np * k + i];
      dv += (double)part[2 * (size_t)np * k + np + i];
    }
    pu[i] = (float)((double)pu[i] + du);
    pv[i] = (float)((double)pv[i] + dv);
  }
  return O2D_OK;
}
