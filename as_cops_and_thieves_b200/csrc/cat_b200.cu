// cat_b200.cu — hand-written sm_100a kernels + C ABI for the batched cops-and-thieves step.
//
// One warp owns one world for one lockstep transition (persistent CTAs loop over worlds):
//   state record (coalesced) -> termination test -> action impulses -> 90-ray sensor sweep per
//   agent (the wall edges, staged in shared memory by one TMA bulk copy per CTA, are RASTERISED into
//   a per-agent 1-D depth buffer: lanes = candidate edges, then lanes = (edge, ray) pairs) ->
//   float16 observation chain -> rewards -> team-shared merge -> rigid-body step (position integrate,
//   circle-hull / circle-circle contacts, warm start, sequential impulses) -> auto-reset (Philox) ->
//   outputs + state record (coalesced).  Plus the MAPPO GAE / advantage-normalisation kernels.
//
// Restates /root/reference/src/environments/base_env.py:354-413,521-554,286-352,
// /root/reference/src/agents/entity.py:126-241, cop.py:49-75, thief.py:48-69,
// /root/reference/src/environments/observation_spaces.py:67-131 and the Chipmunk2D 7.0.3
// routines they call (cpSpaceStep, cpArbiterPreStep/ApplyImpulse, CircleToPoly, CircleToCircle,
// cpShapeSegmentQuery, cpPolyShapeSegmentQuery, CircleSegmentQuery, cpBBSegmentQuery) in fp32,
// origin-relative so that ray distances keep ~1e-5 absolute accuracy at coordinates ~1e3.
//
// No CPU fallback, no Triton, no tensor cores (the step is instruction-issue bound, GAE is HBM bound).

#include <cuda.h>            // CUtensorMap types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked
#include <cudaTypedefs.h>    // PFN_cuTensorMapEncodeTiled
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/cat_b200.h"
#include "../../include/cat_philox.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(CAT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
  } while (0)

#include "ray_lists.h"
#include "world_kernel.cuh"
#include "state_view.cuh"
#include "gae_kernels.cuh"

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Every ABI entry runs with the environment's device current and restores the caller's device on return.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
    if (prev == device) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define DEVICE_SCOPE(env)                                                                    \
  DeviceGuard _guard((env)->device);                                                          \
  if (!_guard.ok) return fail(CAT_ERR_CUDA, "cudaSetDevice(" + std::to_string((env)->device) + ") failed")

}  // namespace

// ------------------------------------------------------------------ host side / C ABI
typedef void (*WorldKernel)(const KParams);
struct LaunchShape { int threads, smem, grid; };

// The (cell, ray) candidate lists depend on the map and the sensor parameters only: environments created on the same
// device for the same map share ONE copy (reference-counted; built and uploaded by the first, freed with the last) — ten
// environments on agh-map hold 120 MB of lists, not 1.2 GB, and creating the second one skips the second-long build.
struct SharedRayLists {
  std::vector<unsigned char> key;   // device, sensor parameters, grid rule and every map array the builder reads
  uint4* slots_dev = nullptr;
  uint32_t* ovf_dev = nullptr;
  RayListGrid g{};
  size_t slot_words = 0, ovf_words = 0;
  int refs = 0;
};
static std::mutex g_ray_lists_mutex;
static std::vector<SharedRayLists*> g_ray_lists;

static void ray_lists_release(SharedRayLists* sh) {
  if (!sh) return;
  std::lock_guard<std::mutex> lock(g_ray_lists_mutex);
  if (--sh->refs > 0) return;
  g_ray_lists.erase(std::remove(g_ray_lists.begin(), g_ray_lists.end(), sh), g_ray_lists.end());
  if (sh->slots_dev) cudaFree(sh->slots_dev);
  if (sh->ovf_dev) cudaFree(sh->ovf_dev);
  delete sh;
}

struct CatEnv {
  std::vector<std::pair<int, LaunchShape>> shape_cache;   // launch shape per world count (the chunked host path asks every step)
  WorldKernel kernel = nullptr;        // every output the ABI offers
  WorldKernel kernel_record = nullptr; // record output only (launches that ask for nothing else)
  int device = 0;
  int n_worlds = 0;
  unsigned char* blob_dev = nullptr;
  int32_t* view_off_dev = nullptr;
  uint16_t* view_edges_dev = nullptr;
  SharedRayLists* ray_lists = nullptr;   // shared with the other environments of this map on this device
  unsigned long long* overflow_dev = nullptr;
  CatRecordLayout rec{}, rec_packed{};
  KParams kp{};
  CatEnvInfo info{};
  int smem_bytes = 0;
  int grid = 0;
  int threads = kThreads;
  int blob_bytes = 0, max_optin = 0, n_sm = 0;
  // chunked host path: a second stream for the device-to-host copies and one event per chunk
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> chunk_events;
  cudaEvent_t copies_done = nullptr;
};

// the kernel instantiation an environment runs (fixed at creation): agents / rays as compile-time constants for the
// shipped shape, the any-shape instantiation otherwise (CAT_GENERIC_KERNEL=1 at creation forces it: test knob)
static WorldKernel pick_world_kernel(int A, int R, bool lists, bool extras) {
  const char* e = getenv("CAT_GENERIC_KERNEL");
  const bool generic = e && e[0] == '1';
  if (A == 3 && R == 90 && !generic) {
    if (lists) return extras ? cat_world_kernel<3, 90, true, true> : cat_world_kernel<3, 90, true, false>;
    return cat_world_kernel<3, 90, false, true>;
  }
  return cat_world_kernel<0, 0, false, true>;
}

static bool pick_launch_shape(CatEnv* env, int n_worlds, LaunchShape* out) {
  for (const auto& kv : env->shape_cache)
    if (kv.first == n_worlds) { *out = kv.second; return true; }
  long long best_score = -1;
  int wpc_max = kMaxThreads / 32;
  if (const char* e = getenv("CAT_MAX_WARPS_PER_CTA")) { const int v = atoi(e); if (v >= 2 && v < wpc_max) wpc_max = v; }   // tuning knob
  for (int wpc = wpc_max; wpc >= 2; --wpc) {   // any warp count: 4096 worlds = 586 CTAs of 7 warps = 28 warps per SM
    const int smem = align_up(env->blob_bytes, 128) + wpc * env->kp.lay.scratch_bytes;
    if (smem > env->max_optin) continue;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, env->kernel, wpc * 32, smem) != cudaSuccess || occ < 1) continue;
    const int need = (n_worlds + wpc - 1) / wpc;
    const bool one_wave = need <= env->n_sm * occ;
    // one wave: fewest warps on the fullest SM; persistent: most resident warps per SM
    const int key = one_wave ? 4096 - ((need + env->n_sm - 1) / env->n_sm) * wpc : occ * wpc;
    const long long score = ((long long)(one_wave ? 1 : 0) << 40) + ((long long)key << 8) + wpc;
    if (score > best_score) {
      best_score = score;
      out->threads = wpc * 32; out->smem = smem; out->grid = one_wave ? need : env->n_sm * occ;
    }
  }
  if (const char* e = getenv("CAT_MAX_GRID")) { const int v = atoi(e); if (v >= 1 && v < out->grid) out->grid = v; }   // tuning knob
  if (best_score >= 0 && env->shape_cache.size() < 64) env->shape_cache.emplace_back(n_worlds, *out);
  return best_score >= 0;
}

// Launch with programmatic stream serialisation: the kernel's CTAs may be scheduled while the previous kernel in the
// stream drains; the kernels call griddepcontrol.wait before they touch global memory (gae_kernels.cuh).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int allowed = [] { const char* e = getenv("CAT_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
  attr[0].val.programmaticStreamSerializationAllowed = allowed;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

extern "C" {

#ifdef CAT_STATS
int cat_debug_stats(unsigned long long* out8, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, g_stats, sizeof(g_stats));
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_stats, z, sizeof(z)); }
  return 0;
}
#endif

int cat_abi_version(void) { return CAT_ABI_VERSION; }
const char* cat_last_error(void) { return g_last_error.c_str(); }

int cat_env_create(const CatMapDesc* map, const CatParams* pr, int32_t n_worlds, int64_t gid0, int32_t device,
                   CatEnv** out) {
  if (!map || !pr || !out) return fail(CAT_ERR_INVALID, "null argument");
  *out = nullptr;
  const int A = map->n_cops + map->n_thieves, R = pr->n_rays, H = map->n_hulls, E = map->n_edges;
  if (n_worlds < 1) return fail(CAT_ERR_INVALID, "n_worlds must be >= 1");
  if (map->n_cops < 1 || map->n_thieves < 1) return fail(CAT_ERR_INVALID, "need at least one cop and one thief");
  if (A > CAT_MAX_AGENTS) return fail(CAT_ERR_LIMIT, "too many agents (CAT_MAX_AGENTS)");
  if (R < 2 || R > CAT_MAX_RAYS) return fail(CAT_ERR_LIMIT, "n_rays out of range (2..CAT_MAX_RAYS)");
  if (H < 1 || H > 65535 || E > 65535) return fail(CAT_ERR_LIMIT, "hull/edge count out of range");
  const int ncell = map->nx * map->ny;
  if (ncell < 1 || map->con_cell_off[ncell] > 65535)
    return fail(CAT_ERR_LIMIT, "grid lists too long for 16-bit offsets");
  for (int h = 0; h < H; ++h)
    if (map->hull_off[h + 1] - map->hull_off[h] > 65535) return fail(CAT_ERR_LIMIT, "hull too large");
  if (!(pr->dt > 0.0)) return fail(CAT_ERR_INVALID, "dt must be > 0");
  const int P = A * (A - 1) / 2;
  if (P > 32) return fail(CAT_ERR_LIMIT, "too many agent pairs for one warp");

  // ---- blob
  const int nreg = map->region_off[A];
  BlobHeader hd{};
  int off = sizeof(BlobHeader);
  auto take = [&](int bytes) { int o = off; off = align_up(off + bytes, 16); return o; };
  hd.n_hulls = H; hd.n_edges = E; hd.nx = map->nx; hd.ny = map->ny;
  hd.gx0 = (float)map->grid_x0; hd.gy0 = (float)map->grid_y0; hd.cell = (float)map->cell;
  hd.inv_cell = (float)(1.0 / map->cell);
  hd.off_edge = take(E * 16);
  hd.off_len = take(E * 4);
  hd.off_hbb = take(H * 16);
  hd.off_heo = take(H * 4);
  hd.off_nextn = take(E * 8);
  hd.off_edgehull = take(E * 2 + 2);
  hd.off_conoff = take((ncell + 1) * 2);
  hd.off_conlist = take(map->con_cell_off[ncell] * 2 + 2);
  hd.off_dir = take(R * 16);
  hd.off_regoff = take((A + 1) * 4);
  hd.off_regions = take((nreg > 0 ? nreg : 1) * 16);
  hd.off_initpos = take(A * 8);
  hd.off_batchbb = take(((E + 31) / 32) * 16);
  const int blob_bytes = align_up(off, 16);
  std::vector<unsigned char> blob(blob_bytes, 0);
  memcpy(blob.data(), &hd, sizeof(hd));
  {
    float* e = reinterpret_cast<float*>(blob.data() + hd.off_edge);
    float* l = reinterpret_cast<float*>(blob.data() + hd.off_len);
    for (int i = 0; i < E; ++i) {
      e[4 * i + 0] = (float)map->vert[2 * i]; e[4 * i + 1] = (float)map->vert[2 * i + 1];
      e[4 * i + 2] = (float)map->normal[2 * i]; e[4 * i + 3] = (float)map->normal[2 * i + 1];
      l[i] = (float)map->edge_len[i];
    }
    float* bb = reinterpret_cast<float*>(blob.data() + hd.off_hbb);
    uint32_t* eo = reinterpret_cast<uint32_t*>(blob.data() + hd.off_heo);
    for (int h = 0; h < H; ++h) {
      // shape bb = raw hull bb grown by the wall radius; rounded outwards so fp32 never shrinks it
      bb[4 * h + 0] = nextafterf((float)(map->hull_bb[4 * h + 0] - pr->wall_radius), -INFINITY);
      bb[4 * h + 1] = nextafterf((float)(map->hull_bb[4 * h + 1] - pr->wall_radius), -INFINITY);
      bb[4 * h + 2] = nextafterf((float)(map->hull_bb[4 * h + 2] + pr->wall_radius), INFINITY);
      bb[4 * h + 3] = nextafterf((float)(map->hull_bb[4 * h + 3] + pr->wall_radius), INFINITY);
      eo[h] = (uint32_t)map->hull_off[h] | ((uint32_t)(map->hull_off[h + 1] - map->hull_off[h]) << 16);
    }
    uint16_t* co = reinterpret_cast<uint16_t*>(blob.data() + hd.off_conoff);
    for (int c = 0; c <= ncell; ++c) co[c] = (uint16_t)map->con_cell_off[c];
    uint16_t* eh = reinterpret_cast<uint16_t*>(blob.data() + hd.off_edgehull);
    float* nn = reinterpret_cast<float*>(blob.data() + hd.off_nextn);
    for (int h = 0; h < H; ++h) {
      const int o = map->hull_off[h], e = map->hull_off[h + 1];
      for (int i = o; i < e; ++i) {
        const int nx_ = (i + 1 < e) ? i + 1 : o;
        eh[i] = (uint16_t)h;
        nn[2 * i] = (float)map->normal[2 * nx_]; nn[2 * i + 1] = (float)map->normal[2 * nx_ + 1];
      }
    }
    uint16_t* cl = reinterpret_cast<uint16_t*>(blob.data() + hd.off_conlist);
    for (int i = 0; i < map->con_cell_off[ncell]; ++i) cl[i] = (uint16_t)map->con_cell_hulls[i];
    float* dir = reinterpret_cast<float*>(blob.data() + hd.off_dir);
    const double step = (2.0 * M_PI) / (double)R;  // entity.py:182 linspace(0, 2pi, R, endpoint=False)
    for (int i = 0; i < R; ++i) {
      const float ux = (float)cos(i * step), uy = (float)sin(i * step);
      dir[4 * i] = ux; dir[4 * i + 1] = uy;
      dir[4 * i + 2] = ux != 0.f ? 1.f / ux : 0.f; dir[4 * i + 3] = uy != 0.f ? 1.f / uy : 0.f;
    }
    int32_t* rg = reinterpret_cast<int32_t*>(blob.data() + hd.off_regoff);
    for (int a = 0; a <= A; ++a) rg[a] = map->region_off[a];
    float* rr = reinterpret_cast<float*>(blob.data() + hd.off_regions);
    for (int i = 0; i < nreg * 4; ++i) rr[i] = (float)map->regions[i];
    float* bbb = reinterpret_cast<float*>(blob.data() + hd.off_batchbb);
    for (int b = 0; b * 32 < E; ++b) {
      float l = INFINITY, bo = INFINITY, r = -INFINITY, t = -INFINITY;
      for (int i = b * 32; i < E && i < b * 32 + 32; ++i) {
        // edge i spans vert[prev(i)] -> vert[i]; prev = i - 1 within the hull, else the hull's last vertex
        int h = 0;
        while (map->hull_off[h + 1] <= i) ++h;
        const int pi = (i > map->hull_off[h]) ? i - 1 : map->hull_off[h + 1] - 1;
        const int idx[2] = {i, pi};
        for (int q = 0; q < 2; ++q) {
          const float x = (float)map->vert[2 * idx[q]], y = (float)map->vert[2 * idx[q] + 1];
          l = fminf(l, x); r = fmaxf(r, x); bo = fminf(bo, y); t = fmaxf(t, y);
        }
      }
      bbb[4 * b + 0] = nextafterf(l, -INFINITY); bbb[4 * b + 1] = nextafterf(bo, -INFINITY);
      bbb[4 * b + 2] = nextafterf(r, INFINITY); bbb[4 * b + 3] = nextafterf(t, INFINITY);
    }
    float* ip = reinterpret_cast<float*>(blob.data() + hd.off_initpos);
    for (int i = 0; i < 2 * A; ++i) ip[i] = (float)map->init_pos[i];
  }

  CatEnv* env = new CatEnv();
  env->device = device;
  env->n_worlds = n_worlds;
  DeviceGuard guard(device);
  if (!guard.ok) { delete env; return fail(CAT_ERR_CUDA, "cudaSetDevice(" + std::to_string(device) + ") failed (no CUDA device?)"); }
  // every failure below goes through cat_env_destroy: nothing allocated so far is leaked
#define CREATE_TRY(expr, what)                                                                              \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) { cat_env_destroy(env); return fail(CAT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(_e)); } \
  } while (0)
  CREATE_TRY(cudaMalloc(&env->blob_dev, blob_bytes), "cudaMalloc(map blob)");
  CREATE_TRY(cudaMemcpy(env->blob_dev, blob.data(), blob_bytes, cudaMemcpyHostToDevice), "cudaMemcpy(map blob)");
  CREATE_TRY(cudaMalloc(&env->overflow_dev, 2 * sizeof(unsigned long long)), "cudaMalloc(overflow counters)");
  CREATE_TRY(cudaMemset(env->overflow_dev, 0, 2 * sizeof(unsigned long long)), "cudaMemset(overflow counters)");

  KParams& k = env->kp;
  k.blob = env->blob_dev; k.blob_bytes = blob_bytes;
  // per-cell candidate lists (optional; valid only if they were built for at least this sensor range)
  // Worth it only when the full scan is long: the lists live in global memory (two dependent loads per sweep,
  // cold after an L2 flush), and a map of a few dozen edges is scanned in one or two batches anyway.
  if (map->view_cell_off && map->view_cell_edges && E >= 96 &&
      map->view_range >= pr->ray_length + pr->wall_radius + pr->ray_radius) {
    const int n_list = map->view_cell_off[ncell];
    std::vector<uint16_t> ve((size_t)(n_list > 0 ? n_list : 1));
    for (int i = 0; i < n_list; ++i) ve[i] = (uint16_t)map->view_cell_edges[i];
    CREATE_TRY(cudaMalloc(&env->view_off_dev, sizeof(int32_t) * (ncell + 1)), "cudaMalloc(view lists)");
    CREATE_TRY(cudaMalloc(&env->view_edges_dev, sizeof(uint16_t) * ve.size()), "cudaMalloc(view lists)");
    CREATE_TRY(cudaMemcpy(env->view_off_dev, map->view_cell_off, sizeof(int32_t) * (ncell + 1), cudaMemcpyHostToDevice), "cudaMemcpy(view lists)");
    CREATE_TRY(cudaMemcpy(env->view_edges_dev, ve.data(), sizeof(uint16_t) * ve.size(), cudaMemcpyHostToDevice), "cudaMemcpy(view lists)");
    k.view_off = env->view_off_dev; k.view_edges = env->view_edges_dev;
  }
  k.overflow = env->overflow_dev;
  // per-(cell, ray) candidate lists of the sensor sweep (ray_lists.h), built here from the map description
  CatEnvInfo& inf = env->info;
  if (!(pr->ray_list_cell < 0.0)) {
    const double rsum = pr->wall_radius + pr->ray_radius;
    // key: everything build_ray_lists reads
    std::vector<unsigned char> key;
    auto put = [&key](const void* p, size_t n) { const unsigned char* b = static_cast<const unsigned char*>(p); key.insert(key.end(), b, b + n); };
    const char* knob_cells = getenv("CAT_RAY_LIST_CELLS");
    const char* knob_min = getenv("CAT_RAY_LIST_MIN_CELL");
    const double head[6] = {(double)device, (double)R, pr->ray_length, rsum, pr->ray_list_cell,
                            (knob_cells ? atof(knob_cells) : 0.0) * 4096.0 + (knob_min ? atof(knob_min) : 0.0)};
    put(head, sizeof(head));
    put(&H, sizeof(H)); put(&E, sizeof(E));
    put(map->hull_off, sizeof(int32_t) * (H + 1));
    put(map->vert, sizeof(double) * 2 * E);
    put(map->normal, sizeof(double) * 2 * E);
    put(map->hull_bb, sizeof(double) * 4 * H);
    SharedRayLists* sh = nullptr;
    {
      std::lock_guard<std::mutex> lock(g_ray_lists_mutex);
      for (SharedRayLists* c : g_ray_lists)
        if (c->key == key) { sh = c; break; }
      if (!sh) {
        RayLists rl;
        build_ray_lists(map, R, pr->ray_length, rsum, pr->ray_list_cell, &rl);
        sh = new SharedRayLists();
        sh->key.swap(key);
        sh->g = rl.g; sh->slot_words = rl.slots.size(); sh->ovf_words = rl.ovf.size();
        cudaError_t e1 = cudaMalloc(&sh->slots_dev, rl.slots.size() * 4);
        cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc(&sh->ovf_dev, rl.ovf.size() * 4) : e1;
        if (e2 == cudaSuccess) e2 = cudaMemcpy(sh->slots_dev, rl.slots.data(), rl.slots.size() * 4, cudaMemcpyHostToDevice);
        if (e2 == cudaSuccess) e2 = cudaMemcpy(sh->ovf_dev, rl.ovf.data(), rl.ovf.size() * 4, cudaMemcpyHostToDevice);
        if (e2 != cudaSuccess) {
          if (sh->slots_dev) cudaFree(sh->slots_dev);
          if (sh->ovf_dev) cudaFree(sh->ovf_dev);
          delete sh;
          cat_env_destroy(env);
          return fail(CAT_ERR_CUDA, std::string("ray lists: ") + cudaGetErrorString(e2));
        }
        g_ray_lists.push_back(sh);
      }
      ++sh->refs;
    }
    env->ray_lists = sh;
    k.ray_slots = sh->slots_dev; k.ray_ovf = sh->ovf_dev;
    k.ray_slot_count = (unsigned)(sh->slot_words / 4); k.ray_ovf_words = (unsigned)sh->ovf_words;
    k.rg_x0 = sh->g.x0; k.rg_y0 = sh->g.y0; k.rg_inv_cell = sh->g.inv_cell; k.rg_nx = sh->g.nx; k.rg_ny = sh->g.ny;
    inf.ray_list_cells = sh->g.nx * sh->g.ny; inf.ray_list_nx = sh->g.nx; inf.ray_list_ny = sh->g.ny; inf.ray_list_cell = sh->g.cell;
    inf.ray_list_bytes = (int64_t)(sh->slot_words + sh->ovf_words) * 4;
  }
  k.n_worlds = n_worlds; k.gid0 = gid0;
  k.A = A; k.nc = map->n_cops; k.R = R;
  k.n_edges = E;
  k.lay = make_layout(A, R);          // record / scratch / output-record offsets (world_kernel.cuh)
  CatRecordLayout& rc = env->rec;
  rc.off_dist = 0; rc.off_type = k.lay.r_off_type; rc.off_reward = k.lay.r_off_reward;
  rc.off_terminated = k.lay.r_off_flags; rc.off_truncated = k.lay.r_off_flags + 1; rc.off_winner = k.lay.r_off_flags + 2;
  rc.bytes = k.lay.r_bytes; rc.type_bits = 8;
  CatRecordLayout& rp = env->rec_packed;   // the host-facing form: types packed 4 to a byte
  rp.off_dist = 0; rp.off_type = k.lay.r_off_type; rp.off_reward = k.lay.rp_off_reward;
  rp.off_terminated = k.lay.rp_off_flags; rp.off_truncated = k.lay.rp_off_flags + 1; rp.off_winner = k.lay.rp_off_flags + 2;
  rp.bytes = k.lay.rp_bytes; rp.type_bits = 2;
  k.state_dim = 0;
  for (int a = 0; a < A; ++a) k.state_dim += 4 * R + 2 * (a < map->n_cops ? map->n_cops : map->n_thieves);
  k.dt = (float)pr->dt; k.inv_dt = (float)(1.0 / pr->dt);
  k.impulse = (float)pr->unit_velocity; k.inv_mass = (float)(1.0 / pr->unit_mass);
  k.agent_r = (float)pr->unit_size; k.max_speed = (float)pr->max_speed; k.term_r = (float)pr->termination_radius;
  k.ray_len = (float)pr->ray_length; k.ray_r = (float)pr->ray_radius; k.wall_r = (float)pr->wall_radius;
  k.slop = (float)pr->collision_slop; k.bias_coef = (float)(1.0 - pow(pr->collision_bias, pr->dt));
  k.iterations = pr->iterations; k.persistence = pr->collision_persistence; k.max_steps = pr->max_step_count;
  k.stale = pr->stale_shape_cache; k.auto_reset = pr->auto_reset; k.seed = pr->seed;

  // Launch shape.  A warp owns a world, so with few worlds (one resident wave) the step takes as long as the
  // SM holding the most warps: smaller CTAs spread N warps more evenly over the SMs (4096 worlds = 27.7 warps per
  // SM: 8-warp CTAs put 32 on some SMs, 4-warp CTAs at most 28).  With many worlds the grid is persistent and the
  // shape with the most resident warps wins (ties: the larger CTA, fewer copies of the map).
  int max_optin = 0, n_sm = 0;
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
  const int smem_max = align_up(blob_bytes, 128) + 2 * k.lay.scratch_bytes;
  if (align_up(blob_bytes, 128) + 2 * k.lay.scratch_bytes > max_optin) {
    cat_env_destroy(env);
    return fail(CAT_ERR_LIMIT, "map does not fit in shared memory (" + std::to_string(smem_max) + " B)");
  }
  // The attribute belongs to the FUNCTION (per device), not to this environment: always raise it to the device
  // maximum, so creating an environment for a smaller map never lowers the cap of one that is still alive.
  cudaFuncAttributes fattr{};
  env->kernel = pick_world_kernel(A, R, k.ray_slots != nullptr, true);
  env->kernel_record = pick_world_kernel(A, R, k.ray_slots != nullptr, false);
  CREATE_TRY(cudaFuncGetAttributes(&fattr, env->kernel), "cudaFuncGetAttributes");
  max_optin -= (int)fattr.sharedSizeBytes;      // the opt-in limit covers static + dynamic shared memory
  CREATE_TRY(cudaFuncSetAttribute(env->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin), "cudaFuncSetAttribute");
  if (env->kernel_record != env->kernel)
    CREATE_TRY(cudaFuncSetAttribute(env->kernel_record, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin), "cudaFuncSetAttribute");
  env->blob_bytes = blob_bytes; env->max_optin = max_optin; env->n_sm = n_sm;
  LaunchShape shp;
  if (!pick_launch_shape(env, n_worlds, &shp)) { cat_env_destroy(env); return fail(CAT_ERR_CUDA, "occupancy query failed"); }
  env->threads = shp.threads; env->smem_bytes = shp.smem; env->grid = shp.grid;
  k.world_begin = 0; k.world_end = n_worlds;

  inf.n_worlds = n_worlds; inf.n_agents = A; inf.n_cops = map->n_cops; inf.n_thieves = map->n_thieves;
  inf.n_rays = R; inf.n_hulls = H; inf.n_edges = E; inf.state_dim = k.state_dim; inf.record_words = k.lay.rec_words;
  inf.map_blob_bytes = blob_bytes; inf.smem_bytes_per_cta = env->smem_bytes; inf.warps_per_cta = env->threads / 32;
  inf.grid = env->grid; inf.n_pairs = P;
#undef CREATE_TRY
  *out = env;
  return CAT_OK;
}

int cat_env_destroy(CatEnv* env) {
  if (!env) return CAT_OK;
  DeviceGuard guard(env->device);
  if (env->blob_dev) cudaFree(env->blob_dev);
  if (env->view_off_dev) cudaFree(env->view_off_dev);
  if (env->view_edges_dev) cudaFree(env->view_edges_dev);
  ray_lists_release(env->ray_lists);
  if (env->overflow_dev) cudaFree(env->overflow_dev);
  for (cudaEvent_t e : env->chunk_events) cudaEventDestroy(e);
  if (env->copies_done) cudaEventDestroy(env->copies_done);
  if (env->copy_stream) cudaStreamDestroy(env->copy_stream);
  delete env;
  return CAT_OK;
}

int cat_env_info(const CatEnv* env, CatEnvInfo* info) {
  if (!env || !info) return fail(CAT_ERR_INVALID, "null argument");
  *info = env->info;
  return CAT_OK;
}

int cat_ray_lists_host(const CatMapDesc* map, int32_t n_rays, double ray_length, double rsum, double cell, double grid_out[5],
                       uint32_t* slots, int64_t* n_slot_words, uint32_t* ovf, int64_t* n_ovf_words) {
  if (!map || !grid_out || !n_slot_words || !n_ovf_words) return fail(CAT_ERR_INVALID, "null argument");
  if (n_rays < 2 || n_rays > CAT_MAX_RAYS || map->n_hulls < 1 || map->n_edges > 65535) return fail(CAT_ERR_LIMIT, "ray / hull / edge count out of range");
  RayLists rl;
  build_ray_lists(map, n_rays, ray_length, rsum, cell, &rl);
  grid_out[0] = rl.g.x0; grid_out[1] = rl.g.y0; grid_out[2] = rl.g.cell; grid_out[3] = rl.g.nx; grid_out[4] = rl.g.ny;
  if (slots && ovf) {
    if (*n_slot_words < (int64_t)rl.slots.size() || *n_ovf_words < (int64_t)rl.ovf.size()) return fail(CAT_ERR_INVALID, "buffers too small");
    memcpy(slots, rl.slots.data(), rl.slots.size() * 4);
    memcpy(ovf, rl.ovf.data(), rl.ovf.size() * 4);
  }
  *n_slot_words = (int64_t)rl.slots.size(); *n_ovf_words = (int64_t)rl.ovf.size();
  return CAT_OK;
}

int cat_env_record_layout(const CatEnv* env, CatRecordLayout* layout) {
  if (!env || !layout) return fail(CAT_ERR_INVALID, "null argument");
  *layout = env->rec;
  return CAT_OK;
}

int cat_env_packed_record_layout(const CatEnv* env, CatRecordLayout* layout) {
  if (!env || !layout) return fail(CAT_ERR_INVALID, "null argument");
  *layout = env->rec_packed;
  return CAT_OK;
}

int cat_env_overflow_counts(CatEnv* env, uint64_t out[2], int32_t reset) {
  if (!env || !out) return fail(CAT_ERR_INVALID, "null argument");
  DEVICE_SCOPE(env);
  CUDA_TRY(cudaDeviceSynchronize());
  unsigned long long v[2] = {0, 0};
  CUDA_TRY(cudaMemcpy(v, env->overflow_dev, sizeof(v), cudaMemcpyDeviceToHost));
  out[0] = v[0]; out[1] = v[1];
  if (reset) CUDA_TRY(cudaMemset(env->overflow_dev, 0, sizeof(v)));
  return CAT_OK;
}

int cat_env_set_seed(CatEnv* env, uint64_t seed) {
  if (!env) return fail(CAT_ERR_INVALID, "null env");
  env->kp.seed = seed;
  return CAT_OK;
}

size_t cat_env_state_bytes(const CatEnv* env) {
  return env ? (size_t)env->n_worlds * env->kp.lay.rec_words * 4 : 0;
}

static int prepare(CatEnv* env, void* state_dev, const CatStepIO* io, int mode, KParams* out) {
  if (!env || !state_dev) return fail(CAT_ERR_INVALID, "null env/state");
  KParams k = env->kp;
  k.state = reinterpret_cast<float*>(state_dev);
  k.mode = mode;
  if (io) {
    if (mode == MODE_STEP) {
      if (!io->actions) return fail(CAT_ERR_INVALID, "step needs actions");
      if (io->actions_kind < 0 || io->actions_kind > 3) return fail(CAT_ERR_INVALID, "bad actions_kind");
      k.actions_kind = io->actions_kind;
      if (io->actions_kind == 3) for (int a = 0; a < k.A; ++a) k.actions[a] = reinterpret_cast<const void* const*>(io->actions)[a];
      else k.actions[0] = io->actions;
    }
    k.reset_mask = io->reset_mask;
    k.obs_dist = io->obs_dist; k.obs_type = io->obs_type; k.reward = io->reward;
    k.dist_stride = io->obs_dist_world_stride ? io->obs_dist_world_stride : k.lay.nrays * 2;
    k.type_stride = io->obs_type_world_stride ? io->obs_type_world_stride : k.lay.nrays;
    if (k.dist_stride < k.lay.nrays * 2 || (k.dist_stride & 1) || k.type_stride < k.lay.nrays)
      return fail(CAT_ERR_INVALID, "observation world strides are smaller than one world's block");
    k.obs_vec = ((k.dist_stride | k.type_stride) & 15) == 0 &&
                ((reinterpret_cast<uintptr_t>(io->obs_dist) | reinterpret_cast<uintptr_t>(io->obs_type)) & 15) == 0 &&
                k.dist_stride >= (k.lay.nrays * 2 + 15) / 16 * 16 && k.type_stride >= (k.lay.nrays + 15) / 16 * 16;
    k.terminated = io->terminated;
    k.truncated = io->truncated; k.winner = io->winner; k.shared_dist = io->shared_dist; k.shared_type = io->shared_type;
    k.team_pos = io->team_pos; k.obs_f32 = io->obs_f32; k.state_f32 = io->state_f32; k.hit_point = io->hit_point;
    k.critic_f32 = io->critic_f32; k.obs_bf16 = io->obs_bf16; k.critic_bf16 = io->critic_bf16;
    if (io->record) {
      k.record = reinterpret_cast<unsigned char*>(io->record);
      k.record_packed = io->record_packed_types ? 1 : 0;
      const int rec_bytes = k.record_packed ? k.lay.rp_bytes : k.lay.r_bytes;
      k.record_stride = io->record_world_stride ? io->record_world_stride : rec_bytes;
      if ((reinterpret_cast<uintptr_t>(io->record) & 15) || (k.record_stride & 15) || k.record_stride < rec_bytes)
        return fail(CAT_ERR_INVALID, "record output must be 16-byte aligned with a world stride that is a multiple of 16 and >= CatRecordLayout.bytes");
      k.obs_dist = nullptr; k.obs_type = nullptr; k.reward = nullptr; k.terminated = nullptr; k.truncated = nullptr; k.winner = nullptr;
    }
  } else if (mode == MODE_STEP) {
    return fail(CAT_ERR_INVALID, "step needs a CatStepIO");
  }
  *out = k;
  return CAT_OK;
}

// the instantiation for this launch: the record-only one for a step whose record is all it writes
static WorldKernel kernel_for(const CatEnv* env, const KParams& k) {
  const bool record_only = k.mode == MODE_STEP && k.record && !k.shared_dist && !k.shared_type && !k.team_pos && !k.obs_f32 && !k.state_f32 &&
                           !k.hit_point && !k.critic_f32 && !k.obs_bf16 && !k.critic_bf16;
  return record_only ? env->kernel_record : env->kernel;
}

static int launch(CatEnv* env, void* state_dev, const CatStepIO* io, int mode, void* stream) {
  KParams k;
  const int rc = prepare(env, state_dev, io, mode, &k);
  if (rc != CAT_OK) return rc;
  DEVICE_SCOPE(env);
  CUDA_TRY(launch_pdl(kernel_for(env, k), env->grid, env->threads, (size_t)env->smem_bytes, reinterpret_cast<cudaStream_t>(stream), k));
  return CAT_OK;
}

int cat_env_init_state(CatEnv* env, void* state_dev, void* stream) { return launch(env, state_dev, nullptr, MODE_INIT, stream); }
int cat_env_reset(CatEnv* env, void* state_dev, const CatStepIO* io, void* stream) { return launch(env, state_dev, io, MODE_RESET, stream); }
int cat_env_step(CatEnv* env, void* state_dev, const CatStepIO* io, void* stream) { return launch(env, state_dev, io, MODE_STEP, stream); }
int cat_env_observe(CatEnv* env, void* state_dev, const CatStepIO* io, void* stream) { return launch(env, state_dev, io, MODE_OBSERVE, stream); }

// BaseEnv.step for a caller whose buffers are pinned HOST memory, pipelined: the worlds are stepped in
// `n_chunks` consecutive launches on `stream`, each writing its worlds' output RECORDS into `records_dev`; as
// soon as a chunk's launch has finished, its block of records moves to `records_host` with ONE cudaMemcpyAsync on
// an internal copy stream (DMA engine, full PCIe rate) while the next chunk computes.  `stream` finally waits for
// the copies, so one cudaStreamSynchronize(stream) by the caller makes every result visible.
int cat_env_step_host(CatEnv* env, void* state_dev, const uint8_t* host_actions, void* records_dev, void* records_host,
                      int32_t record_world_stride, int32_t packed_types, int32_t n_chunks, void* stream_) {
  if (!env || !host_actions || !records_dev || !records_host) return fail(CAT_ERR_INVALID, "null argument");
  const int N = env->n_worlds;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > N) n_chunks = N;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DEVICE_SCOPE(env);
  if (!env->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&env->copy_stream, cudaStreamNonBlocking));
  if (!env->copies_done) CUDA_TRY(cudaEventCreateWithFlags(&env->copies_done, cudaEventDisableTiming));
  while ((int)env->chunk_events.size() < n_chunks) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    env->chunk_events.push_back(e);
  }
  CatStepIO io{};
  io.actions = host_actions; io.actions_kind = 0;
  io.record = records_dev; io.record_world_stride = record_world_stride; io.record_packed_types = packed_types;
  KParams k;
  const int rc = prepare(env, state_dev, &io, MODE_STEP, &k);
  if (rc != CAT_OK) return rc;
  const size_t stride = (size_t)k.record_stride;
  const int per = ((N + n_chunks - 1) / n_chunks + 7) / 8 * 8;   // worlds per chunk, a multiple of the CTA's 8 warps
  int c = 0;
  for (int w0 = 0; w0 < N; w0 += per, ++c) {
    const int w1 = w0 + per < N ? w0 + per : N, n = w1 - w0;
    LaunchShape shp;
    if (!pick_launch_shape(env, n, &shp)) return fail(CAT_ERR_CUDA, "occupancy query failed");
    k.world_begin = w0; k.world_end = w1;
    CUDA_TRY(launch_pdl(kernel_for(env, k), shp.grid, shp.threads, (size_t)shp.smem, stream, k));
    CUDA_TRY(cudaEventRecord(env->chunk_events[c], stream));
    CUDA_TRY(cudaStreamWaitEvent(env->copy_stream, env->chunk_events[c], 0));
    CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(records_host) + (size_t)w0 * stride, static_cast<const char*>(records_dev) + (size_t)w0 * stride,
                             (size_t)n * stride, cudaMemcpyDeviceToHost, env->copy_stream));
  }
  CUDA_TRY(cudaEventRecord(env->copies_done, env->copy_stream));
  CUDA_TRY(cudaStreamWaitEvent(stream, env->copies_done, 0));
  return CAT_OK;
}

static int state_view(CatEnv* env, void* state_dev, const CatStateView* view, int set, void* stream) {
  if (!env || !state_dev || !view) return fail(CAT_ERR_INVALID, "null argument");
  const KParams& k = env->kp;
  ViewParams p{};
  p.state = reinterpret_cast<float*>(state_dev);
  p.rec_words = k.lay.rec_words; p.n_worlds = k.n_worlds; p.A = k.A; p.P = k.lay.P;
  p.o_vel = k.lay.o_vel; p.o_vb = k.lay.o_vb; p.o_tc = k.lay.o_tc; p.o_wkey = k.lay.o_wkey; p.o_wjn = k.lay.o_wjn;
  p.o_page = k.lay.o_page; p.o_pjn = k.lay.o_pjn; p.o_sc = k.lay.o_sc; p.o_ep = k.lay.o_ep; p.o_flags = k.lay.o_flags;
  p.v = *view; p.set = set;
  DEVICE_SCOPE(env);
  const int threads = 128, blocks = (k.n_worlds + threads - 1) / threads;
  cat_state_view_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  CUDA_TRY(cudaGetLastError());
  return CAT_OK;
}

int cat_env_get_state(CatEnv* env, const void* state_dev, const CatStateView* view, void* stream) {
  return state_view(env, const_cast<void*>(state_dev), view, 0, stream);
}
int cat_env_set_state(CatEnv* env, void* state_dev, const CatStateView* view, void* stream) {
  return state_view(env, state_dev, view, 1, stream);
}

// ---- tensor maps for the TMA-fed GAE kernel
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static bool gae_tma_usable(const void* r, const void* d, const void* v, int M) {
  const char* env = getenv("CAT_GAE_TMA");   // opt-out (A/B against the register-pipelined kernel); a libc table look-up, ~50 ns
  if (env && env[0] == '0') return false;
  if (M % 16 != 0) return false;                                           // u8 row stride must be a multiple of 16 bytes
  if ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(v)) & 15) return false;
  return tensor_map_encoder() != nullptr;
}

static bool encode_2d(CUtensorMap* tm, CUtensorMapDataType dt, const void* ptr, int M, int T, int elem) {
  const cuuint64_t dims[2] = {(cuuint64_t)M, (cuuint64_t)T};
  const cuuint64_t strides[1] = {(cuuint64_t)M * elem};
  const cuuint32_t box[2] = {(cuuint32_t)kTmaCols, (cuuint32_t)kTmaRows};
  const cuuint32_t estr[2] = {1, 1};
  return tensor_map_encoder()(tm, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int cat_gae(const float* rewards, const uint8_t* dones, const float* values, const float* last_values, float* returns,
            float* advantages, double* stats_dev, int32_t T, int32_t M, float gamma, float lam, void* stream) {
  if (!rewards || !dones || !values || !last_values || !returns || !advantages || !stats_dev)
    return fail(CAT_ERR_INVALID, "null argument");
  if (T < 1 || M < 1) return fail(CAT_ERR_INVALID, "T and M must be >= 1");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int threads = kGaeCols * kGaeSegs, blocks = (M + kGaeCols - 1) / kGaeCols;
  const bool idx32 = (long long)(T + kTmaRows) * M < (1ll << 31);
  if (gae_tma_usable(rewards, dones, values, M)) {
    // TMA-fed variant: tensor maps over (M, T) for r, V (fp32) and done (u8), box 64 columns x 32 steps
    CUtensorMap tm_r, tm_v, tm_d;
    if (encode_2d(&tm_r, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rewards, M, T, 4) && encode_2d(&tm_v, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, values, M, T, 4) &&
        encode_2d(&tm_d, CU_TENSOR_MAP_DATA_TYPE_UINT8, dones, M, T, 1)) {
      const int smem = kTmaStages * kTmaStageBytes, tblocks = (M + kTmaCols - 1) / kTmaCols;
      // the shared-memory opt-in is a per-function, per-device attribute: set once per device, not per call
      static bool attr_set[64] = {false};
      int devid = 0;
      CUDA_TRY(cudaGetDevice(&devid));
      if (devid < 0 || devid >= 64 || !attr_set[devid]) {
        CUDA_TRY(cudaFuncSetAttribute(cat_gae_tma_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CUDA_TRY(cudaFuncSetAttribute(cat_gae_tma_kernel<size_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (devid >= 0 && devid < 64) attr_set[devid] = true;
      }
      if (idx32)
        CUDA_TRY(launch_pdl(cat_gae_tma_kernel<uint32_t>, tblocks, kTmaCols * kTmaSegs, smem, s, tm_r, tm_v, tm_d, last_values, returns, advantages, stats_dev, T, M, gamma, lam));
      else
        CUDA_TRY(launch_pdl(cat_gae_tma_kernel<size_t>, tblocks, kTmaCols * kTmaSegs, smem, s, tm_r, tm_v, tm_d, last_values, returns, advantages, stats_dev, T, M, gamma, lam));
      CUDA_TRY(cudaGetLastError());
      return CAT_OK;
    }
  }
  if (idx32)
    CUDA_TRY(launch_pdl(cat_gae_kernel<uint32_t>, blocks, threads, 0, s, rewards, dones, values, last_values, returns, advantages, stats_dev, T, M, gamma, lam));
  else
    CUDA_TRY(launch_pdl(cat_gae_kernel<size_t>, blocks, threads, 0, s, rewards, dones, values, last_values, returns, advantages, stats_dev, T, M, gamma, lam));
  CUDA_TRY(cudaGetLastError());
  return CAT_OK;
}

int cat_adv_normalize(float* advantages, int64_t n, const double* stats_dev, int64_t count, void* stream) {
  if (!advantages || !stats_dev) return fail(CAT_ERR_INVALID, "null argument");
  if (n < 1 || count < 1) return fail(CAT_ERR_INVALID, "n and count must be >= 1");
  if ((reinterpret_cast<uintptr_t>(advantages) & 15) != 0) return fail(CAT_ERR_INVALID, "advantages must be 16-byte aligned");
  const int threads = 256;
  long long blocks = (n / 4 + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  CUDA_TRY(launch_pdl(cat_adv_normalize_kernel, (int)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream), advantages, (long long)n, stats_dev, (long long)count));
  CUDA_TRY(cudaGetLastError());
  return CAT_OK;
}

}  // extern "C"
