/* CPU oracle for the cops-and-thieves environment step — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product path (as_cops_and_thieves_b200) never links, imports or calls it.
 *
 * PARITY UNPINNED: the arithmetic the reference runs lives in third-party pymunk (unpinned in
 * /root/reference/requirements.txt:4; API style matches pymunk 6.x = Chipmunk2D 7.0.3), which is
 * neither vendored in /root/reference nor installable here, and the reference ships no tests or
 * golden vectors.  This file restates (a) the reference's own Python step logic, each function
 * citing the file:line it follows, and (b) the published Chipmunk2D 7.0.3 algorithms those lines
 * call (cpSpaceStep.c, cpArbiter.c, cpCollision.c, cpPolyShape.c, cpShape.c, cpSpaceQuery.c,
 * cpBBTree.c, cpBB.h), in fp64 like Chipmunk's cpFloat.  It is pinned only by the analytic
 * known-answer vectors in tests/golden/ (authored from first principles).
 */
#ifndef CAT_ORACLE_H
#define CAT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t n_hulls, n_edges;
  const int32_t* hull_off; /* [H+1] */
  const double* vert;      /* [E][2] hull vertices, CCW per hull */
  int32_t n_cops, n_thieves;
  const double* init_pos;    /* [A][2] cops first then thieves */
  const int32_t* region_off; /* [A+1] */
  const double* regions;     /* [R][4] x,y,w,h */
} OrcMap;

typedef struct {
  double dt;
  int32_t max_step_count;
  double unit_velocity, unit_mass, unit_size, max_speed, termination_radius;
  double ray_length, ray_radius, wall_radius;
  int32_t n_rays;
  int32_t iterations;
  double collision_slop, collision_bias;
  int32_t collision_persistence;
  int32_t stale_shape_cache; /* 1 = pymunk behaviour: reset leaves query centres stale */
  int32_t auto_reset;        /* 1 = batched-env behaviour */
  uint64_t seed;
} OrcParams;

typedef struct {
  int32_t n_worlds;
  int64_t gid0; /* global id of world 0 (sharding) */
  double *pos, *vel, *vbias, *tc; /* [N][A][2] */
  int32_t* step_count;            /* [N] */
  uint32_t* episode;              /* [N] */
  double* wall_jn;                /* [N][A][H] accumulated normal impulse of the cached arbiter */
  int8_t* wall_age;               /* [N][A][H] -1 none; k>=0: arbiter last used k steps before the last completed one */
  double* pair_jn;                /* [N][A][A] (i<j used) */
  int8_t* pair_age;               /* [N][A][A] */
} OrcState;

typedef struct { /* any pointer may be NULL */
  uint16_t* obs_dist;   /* f16 bits [N][A][R] */
  uint8_t* obs_type;    /* [N][A][R] */
  double* hit_point;    /* [N][A][R][2] pre-quantisation hit point (ray end when nothing hit) */
  double* hit_alpha;    /* [N][A][R] 1.0 when nothing hit */
  float* reward;        /* [N][A] */
  uint8_t* terminated;  /* [N] captured or timed out (entity.py:146) */
  uint8_t* truncated;   /* [N] timed out (base_env.py:397) */
  int8_t* winner;       /* [N] -1 none, 0 cop, 1 thief (base_env.py:399-411) */
  uint16_t* shared_dist; /* f16 bits [N][2][R] team 0 = cops, 1 = thieves */
  uint8_t* shared_type;  /* [N][2][R] */
  uint16_t* team_pos;    /* f16 bits [N][A][2] */
} OrcOut;

typedef struct OrcEnv OrcEnv;

OrcEnv* orc_create(const OrcMap* map, const OrcParams* params);
void orc_destroy(OrcEnv* env);
int orc_num_threads(void);
void orc_init_state(const OrcEnv* env, OrcState* st);
/* reference reset(): base_env.py:286-352.  mask NULL = all worlds. */
void orc_reset(const OrcEnv* env, OrcState* st, const uint8_t* mask, OrcOut* out);
/* reference step(): base_env.py:354-413.  actions int32 [N][A] in {0..3}. */
void orc_step(const OrcEnv* env, OrcState* st, const int32_t* actions, OrcOut* out);
/* observation of the current state without stepping (entity.py:159-220 + observation_spaces.py:67-131) */
void orc_observe(const OrcEnv* env, const OrcState* st, OrcOut* out);

/* skrl MAPPO GAE + advantage normalisation, fp64 accumulations.
 * rewards/values/dones are [T][M]; last_values [M]; outputs [T][M]. */
void orc_gae(const float* rewards, const uint8_t* dones, const float* values, const float* last_values,
             float* returns, float* advantages, int T, int M, double gamma, double lam, int normalize);

/* helpers exposed for unit tests */
uint16_t orc_double_to_half_bits(double x);
float orc_half_bits_to_float(uint16_t h);
/* nearest hit of a fat ray against one world's shapes; returns shape id (-1 none; <H hull; H+j agent j) */
int orc_segment_query_first(const OrcEnv* env, const double* tc /*[A][2]*/, int self_agent /* -1: walls only */,
                            double ax, double ay, double bx, double by, double radius,
                            double* alpha, double* point /*[2]*/);
/* cpSpaceStep alone for every world (no actions, no observations) */
void orc_space_step(const OrcEnv* env, OrcState* st);
/* nearest shape within maxd of a point (own shape excluded); returns shape id like orc_segment_query_first */
int orc_point_query_nearest(const OrcEnv* env, const double* tc, int self_agent, double px, double py, double maxd,
                            double* dist_out);
/* signed distance from p to hull h (negative inside), raw hull (no radius) */
double orc_hull_distance(const OrcEnv* env, int h, double px, double py);
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out4);

#ifdef __cplusplus
}
#endif
#endif
