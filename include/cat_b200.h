/* cat_b200 — C ABI of the B200-native batched cops-and-thieves environment step.
 *
 * The reference (Hevagog/as-cops-and-thieves) is pure Python over pymunk's CFFI; it has no FFI
 * seam of its own.  Each entry point below replaces the reference call(s) named beside it; the
 * Python class that mirrors the reference's PettingZoo / skrl surface
 * (as_cops_and_thieves_b200/env.py) is the only caller.  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every `*_dev` / device pointer is CUDA device memory owned by the caller
 *     (PyTorch tensors); the library owns only the staged map constants inside CatEnv;
 *   - all work is enqueued on the caller's stream (`stream` = cudaStream_t as void*), no internal
 *     synchronisation, CUDA-graph capturable; one CatEnv per GPU / rank; not thread-safe per CatEnv;
 *   - return 0 on success, negative CatStatus on failure; never throws, never exits;
 *     cat_last_error() returns a thread-local message for the last failure;
 *   - there is NO CPU fallback: every compute entry point fails with CAT_ERR_CUDA without a GPU.
 *
 * Layouts (N worlds, A = n_cops + n_thieves agents ordered cops then thieves
 * [base_env.py:91-96], R rays):
 *   obs_dist  f16 [N][A][R]   entity.py:200-210   (float16 chain reproduced bit-for-bit)
 *   obs_type  u8  [N][A][R]   entity.py:222-241   (WALL 0, COP 1, THIEF 2, EMPTY 4)
 *   reward    f32 [N][A]      cop.py:49-75, thief.py:48-69
 *   terminated u8 [N]         captured or timed out (entity.py:146)
 *   truncated  u8 [N]         timed out (base_env.py:397)
 *   winner     i8 [N]         -1 none, 0 cop, 1 thief (base_env.py:399-411)
 *   shared_dist f16 [N][2][R], shared_type u8 [N][2][R]   observation_spaces.py:97-121 (team 0 cops, 1 thieves)
 *   team_pos  f16 [N][A][2]   observation_spaces.py:92-95
 *   obs_f32   f32 [A][N][2R]  per-agent obs as skrl flattens it: [distance | object_type]
 *   state_f32 f32 [N][S]      env.state() as skrl flattens it, S = sum_a (4R + 2*team_size(a))
 *   hit_point f32 [N][A][R][2] pre-quantisation hit point (parity/debug)
 */
#ifndef CAT_B200_H
#define CAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAT_ABI_VERSION 5
#define CAT_MAX_AGENTS 8
#define CAT_MAX_RAYS 128
#define CAT_WALL_SLOTS 4 /* cached wall arbiters / simultaneous wall contacts kept per agent; a 5th is counted by
                            cat_env_overflow_counts, never dropped silently */
#define CAT_NEAR_SLOTS 4 /* hulls whose reach can contain a ray origin (alpha = 0 rule) kept per agent; same rule */

typedef enum {
  CAT_OK = 0,
  CAT_ERR_INVALID = -1, /* bad argument */
  CAT_ERR_CUDA = -2,    /* CUDA runtime error (incl. no device) */
  CAT_ERR_LIMIT = -3,   /* map / agent / ray count exceeds a compiled limit */
} CatStatus;

/* Host-side description of one compiled map (what maps.compile_map() yields; replaces
 * Map.populate_space + pymunk.Poly hulls, /root/reference/src/maps/map.py:119-128). */
typedef struct {
  int32_t n_hulls, n_edges;
  const int32_t* hull_off; /* [H+1] */
  const double* vert;      /* [E][2] hull vertices, CCW; edge i runs vert[i-1] -> vert[i] */
  const double* normal;    /* [E][2] outward unit normals */
  const double* edge_len;  /* [E] */
  const double* hull_bb;   /* [H][4] l,b,r,t of the raw hull */
  int32_t n_cops, n_thieves;
  const double* init_pos;    /* [A][2] */
  const int32_t* region_off; /* [A+1] */
  const double* regions;     /* [n_regions][4] x,y,w,h */
  double grid_x0, grid_y0, cell; /* uniform grid used by contact / spawn / start-inside lookups */
  int32_t nx, ny;
  const int32_t* con_cell_off;   /* [nx*ny+1] hulls within contact reach of each cell */
  const int32_t* con_cell_hulls;
  /* Optional (NULL = scan every edge): per grid cell, the edges that can be sensor candidates for some origin
   * inside the cell (facing it and within view_range), nearest first — maps.view_lists().  A conservative
   * superset: the kernel still applies the exact per-origin test to every listed edge, so results do not
   * depend on these lists; they only shorten the candidate scan.  Ignored unless
   * view_range >= ray_length + wall_radius + ray_radius. */
  const int32_t* view_cell_off;   /* [nx*ny+1] */
  const int32_t* view_cell_edges; /* edge ids (< 65535) */
  double view_range;
} CatMapDesc;

/* pyproject.toml:12-19 [tool.physical-params], entity.py:84-86, Chipmunk space defaults, SimpleEnv defaults */
typedef struct {
  double dt;              /* simple_env.py:20 (1/60) */
  int32_t max_step_count; /* simple_env.py:19 (400) */
  double unit_velocity, unit_mass, unit_size, max_speed, termination_radius;
  double ray_length, ray_radius, wall_radius;
  int32_t n_rays;
  int32_t iterations; /* cpSpace iterations (10) */
  double collision_slop, collision_bias;
  int32_t collision_persistence;
  int32_t stale_shape_cache; /* 1 = pymunk behaviour: reset leaves the query centres of the shapes stale */
  int32_t auto_reset;        /* 1 = done worlds are re-spawned inside the step and emit the reset observation */
  uint64_t seed;
  /* Sensor-sweep acceleration (never changes a result, tests/test_parity_gpu.py::test_candidate_lists_do_not_change_results):
   * the library builds, per cell of a uniform grid over the walls' reach and per ray index, the list of edges that
   * ray can touch from any origin in the cell, nearest first.  ray_list_cell = edge length of those cells in map
   * units; 0 = choose automatically (about 65536 cells, not below 6 units: 60-120 MB of lists); < 0 = no lists: every sweep rasterises
   * the map's edges into the per-agent depth buffer (the slower, any-map path). */
  double ray_list_cell;
} CatParams;

typedef struct {
  /* input (step only) */
  const void* actions;   /* kind 0: u8 [N][A]; 1: i32 [N][A]; 2: i64 [N][A]; 3: HOST array of A device pointers to i64 [N] */
  int32_t actions_kind;
  /* input (reset only): u8 [N] device mask, NULL = every world */
  const uint8_t* reset_mask;
  /* outputs — device pointers, any may be NULL */
  uint16_t* obs_dist;
  uint8_t* obs_type;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  int8_t* winner;
  uint16_t* shared_dist;
  uint8_t* shared_type;
  uint16_t* team_pos;
  float* obs_f32;
  float* state_f32;
  float* hit_point;
  /* bytes between consecutive worlds in obs_dist / obs_type; 0 = dense (A*R*2 and A*R).  A stride that is a
   * multiple of 16 (with a 16-byte aligned base) lets the kernel store each world's observation with 16-byte
   * vector stores (the world's block is then written up to the next multiple of 16 bytes, padding included). */
  int32_t obs_dist_world_stride;
  int32_t obs_type_world_stride;
  /* RECORD OUTPUT (preferred): one contiguous record per world, CatRecordLayout below —
   *   [ f16 distance A*R | pad to 16 | u8 type A*R | pad to 4 | f32 reward A | u8 terminated | u8 truncated | i8 winner | pad to 16 ]
   * written with 16-byte stores (device memory or mapped pinned host memory).  When `record` is non-NULL the six
   * pointers obs_dist, obs_type, reward, terminated, truncated, winner above are ignored: one world = one block,
   * so a host-facing caller moves a range of worlds with ONE copy.  record must be 16-byte aligned and
   * record_world_stride a multiple of 16 that is >= CatRecordLayout.bytes (0 = exactly that). */
  void* record;
  int32_t record_world_stride;
  /* critic front end (lstm_value_net.py:122-137): the 4 ray channels LSTMValue cuts from the first agent's block of
   * env.state(), in the order it stacks them — [own_obj_types | own_distances | object_type_shared | distance_shared]
   * of cop_0 — f32 [N][4][R] */
  float* critic_f32;
  /* bf16 copies of obs_f32 ([A][N][2R]) and critic_f32 ([N][4][R]) for mixed-precision learners (optional) */
  uint16_t* obs_bf16;
  uint16_t* critic_bf16;
  /* != 0: `record` takes the PACKED form of the record (cat_env_packed_record_layout; types at 2 bits each, 640 B
   * instead of 832 B per world for 3 agents x 90 rays) — the form meant to cross PCIe on the host-facing path. */
  int32_t record_packed_types;
} CatStepIO;

/* byte offsets inside one world's output record (cat_env_record_layout / cat_env_packed_record_layout) */
typedef struct {
  int32_t bytes;        /* record size = default world stride (multiple of 16) */
  int32_t off_dist;     /* f16 [A][R] */
  int32_t off_type;     /* type_bits == 8: u8 [A][R], ObjectType values (0 wall, 1 cop, 2 thief, 4 empty);
                         * type_bits == 2: ray r = a * R + i of the world is bits 2 (r % 4) .. + 1 of byte r / 4,
                         *   codes 0 wall, 1 cop, 2 thief, 3 empty */
  int32_t off_reward;   /* f32 [A] */
  int32_t off_terminated, off_truncated, off_winner; /* u8, u8, i8 */
  int32_t type_bits;    /* 8 or 2 */
} CatRecordLayout;

typedef struct {
  int32_t n_worlds, n_agents, n_cops, n_thieves, n_rays, n_hulls, n_edges;
  int32_t state_dim;          /* S of state_f32 */
  int32_t record_words;       /* 4-byte words per world in the packed state buffer */
  int32_t map_blob_bytes;     /* bytes staged into shared memory per CTA */
  int32_t smem_bytes_per_cta;
  int32_t warps_per_cta;
  int32_t grid;               /* CTAs launched by cat_env_step */
  int32_t n_pairs;
  int32_t ray_list_cells;     /* cells of the (cell, ray) candidate-list grid, 0 = lists disabled */
  int32_t ray_list_nx, ray_list_ny;
  float ray_list_cell;
  int64_t ray_list_bytes;     /* device bytes held by the lists (slots + overflow) */
} CatEnvInfo;

/* SoA view of the world state for get/set (device pointers; any may be NULL = skip). */
typedef struct {
  float* pos;          /* [N][A][2] */
  float* vel;          /* [N][A][2] */
  float* vbias;        /* [N][A][2] */
  float* tc;           /* [N][A][2] cached shape centres the queries see */
  int32_t* step_count; /* [N] */
  uint32_t* episode;   /* [N] */
  int32_t* wall_hull;  /* [N][A][CAT_WALL_SLOTS] hull id, -1 empty */
  int32_t* wall_age;   /* [N][A][CAT_WALL_SLOTS] */
  float* wall_jn;      /* [N][A][CAT_WALL_SLOTS] */
  int32_t* pair_age;   /* [N][P] pairs (i<j) in i-major order, -1 empty */
  float* pair_jn;      /* [N][P] */
} CatStateView;

typedef struct CatEnv CatEnv;

int cat_abi_version(void);
const char* cat_last_error(void);

/* BaseEnv.__init__ (base_env.py:51-121): stage the map, fix the constants.  gid0 = global id of
 * this rank's world 0 (spawn RNG is keyed by global world id). */
int cat_env_create(const CatMapDesc* map, const CatParams* params, int32_t n_worlds, int64_t gid0,
                   int32_t device, CatEnv** out);
int cat_env_destroy(CatEnv* env);
int cat_env_info(const CatEnv* env, CatEnvInfo* info);
int cat_env_record_layout(const CatEnv* env, CatRecordLayout* layout);
/* the record with CatStepIO.record_packed_types != 0 (what cat_env_step_host moves when packed_types != 0) */
int cat_env_packed_record_layout(const CatEnv* env, CatRecordLayout* layout);
/* Fixed-capacity bookkeeping that the reference (Chipmunk) does not have: out[0] = wall contacts beyond
 * CAT_WALL_SLOTS per agent, out[1] = near hulls beyond CAT_NEAR_SLOTS per agent, summed over every launch since
 * creation (or since the last call with reset != 0).  Synchronises the device.  Zero means no world ever diverged
 * from the uncapped algorithm for this reason. */
int cat_env_overflow_counts(CatEnv* env, uint64_t out[2], int32_t reset);
/* reset(seed=...) (base_env.py:307-311): re-key the spawn RNG of this environment */
int cat_env_set_seed(CatEnv* env, uint64_t seed);
/* bytes of the caller-owned packed state buffer (`state_dev` below) */
size_t cat_env_state_bytes(const CatEnv* env);

/* fresh-environment state: agents at their map positions (entity.py:115), nothing cached */
int cat_env_init_state(CatEnv* env, void* state_dev, void* stream);
/* BaseEnv.reset (base_env.py:286-352) for the masked worlds */
int cat_env_reset(CatEnv* env, void* state_dev, const CatStepIO* io, void* stream);
/* BaseEnv.step (base_env.py:354-413) for every world, one launch */
int cat_env_step(CatEnv* env, void* state_dev, const CatStepIO* io, void* stream);
/* BaseEnv.step for a caller whose buffers live in pinned HOST memory (the reference's own calling convention:
 * numpy in, numpy out), pipelined: the worlds are stepped in n_chunks consecutive launches on `stream` that write
 * record output (CatRecordLayout, stride record_world_stride) into `records_dev`; as soon as a chunk's launch has
 * finished its block of records is moved to `records_host` (pinned) with ONE cudaMemcpyAsync on an internal copy
 * stream while the next chunk computes.  host_actions: u8 [N][A] in pinned host memory, read by the kernel
 * directly.  `stream` waits for the copies: one cudaStreamSynchronize(stream) makes every result visible.
 * packed_types != 0: the records have the packed form (cat_env_packed_record_layout) — 23 % fewer bytes over PCIe. */
int cat_env_step_host(CatEnv* env, void* state_dev, const uint8_t* host_actions, void* records_dev,
                      void* records_host, int32_t record_world_stride, int32_t packed_types, int32_t n_chunks,
                      void* stream);
/* Entity.get_observation + get_shared_observations of the current state (entity.py:159-220,
 * observation_spaces.py:67-131) without stepping */
int cat_env_observe(CatEnv* env, void* state_dev, const CatStepIO* io, void* stream);
/* parity-test access to the hidden state */
int cat_env_get_state(CatEnv* env, const void* state_dev, const CatStateView* view, void* stream);
int cat_env_set_state(CatEnv* env, void* state_dev, const CatStateView* view, void* stream);

/* Inspection / test entry point (host only, needs no GPU): the per-(cell, ray) candidate lists cat_env_create builds for
 * this map and these sensor parameters (csrc/ray_lists.h).  grid_out = {x0, y0, cell, nx, ny}.  Call with slots = ovf =
 * NULL to get the sizes (in 32-bit words) in n_slot_words / n_ovf_words, then again with buffers of those sizes.
 * slot[(cell * n_rays + ray) * 4 .. +4]: entries (bf16 bits of the lower bound of the hit distance << 16 | edge id),
 * 0x7F80FFFF = end, bit 31 = link to a 4-word chunk of ovf (word offset in the low 31 bits). */
int cat_ray_lists_host(const CatMapDesc* map, int32_t n_rays, double ray_length, double rsum, double cell,
                       double grid_out[5], uint32_t* slots, int64_t* n_slot_words, uint32_t* ovf, int64_t* n_ovf_words);

/* skrl MAPPO._update GAE (call site agent_learning_utils.py:198-199).
 * rewards/values [T][M] f32, dones [T][M] u8, last_values [M]; returns/advantages [T][M].
 * stats_dev: CAT_GAE_STATS_DOUBLES (6) doubles = { slot0[2], slot1[2], u64 calls, u64 ticket }, ZEROED BY THE CALLER
 * ONCE when it allocates them and then owned by cat_gae.  After a call, {sum(adv), sum(adv^2)} of that call are in slot
 * (calls & 1), i.e. stats_dev[2 * (calls & 1) .. + 2]; the other slot is zero and takes the next call's sums (so no memset
 * runs in front of the kernel and no CTA waits on a counter at its end: csrc/gae_kernels.cuh).  Calls that share a
 * stats buffer must be ordered (same stream).  Both kernels are launched with programmatic stream serialisation — their
 * prologues overlap the previous kernel's tail (environment CAT_PDL=0 turns that off).
 * Runs the TMA-fed kernel when M % 16 == 0 and the three input arrays are 16-byte aligned, the register-pipelined
 * kernel otherwise (environment CAT_GAE_TMA=0 forces the latter); both give the same results to fp32 rounding. */
#define CAT_GAE_STATS_DOUBLES 6
int cat_gae(const float* rewards, const uint8_t* dones, const float* values, const float* last_values,
            float* returns, float* advantages, double* stats_dev, int32_t T, int32_t M, float gamma,
            float lam, void* stream);
/* advantages = (advantages - mean) / (std + 1e-8) with mean/std (unbiased) from stats over `count` samples.  stats_dev
 * has cat_gae's layout (pass the buffer cat_gae filled).  For a global normalisation across ranks, all-reduce the sums and
 * the count and pass a zeroed CAT_GAE_STATS_DOUBLES buffer with the reduced sums in [0..1] (calls = 0 selects slot 0). */
int cat_adv_normalize(float* advantages, int64_t n, const double* stats_dev, int64_t count, void* stream);

/* Environment variables the library reads (tuning / test knobs; none is needed for normal use, results never depend on them):
 *   CAT_GENERIC_KERNEL=1       cat_env_create: run the any-shape instantiation <0, 0> even for 3 agents x 90 rays (tests)
 *   CAT_MAX_WARPS_PER_CTA=n    pick_launch_shape: cap the warps per CTA (2..32)
 *   CAT_MAX_GRID=n             pick_launch_shape: cap the CTAs of a launch
 *   CAT_RAY_LIST_CELLS=n       automatic list grid: target cell count (default 65536)
 *   CAT_RAY_LIST_MIN_CELL=x    automatic list grid: smallest cell (default 6 units)
 *   CAT_PDL=0                  launch without programmatic stream serialisation (read once per process)
 *   CAT_GAE_TMA=0              cat_gae: use the register-pipelined kernel for every shape
 * Python side: CAT_B200_LIB=path loads another build of this library (profiling / checked builds under variants/). */

#ifdef __cplusplus
}
#endif
#endif /* CAT_B200_H */
