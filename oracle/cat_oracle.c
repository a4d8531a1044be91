/* CPU oracle — TEST INFRASTRUCTURE, NOT PRODUCT.  PARITY UNPINNED (see cat_oracle.h header).
 *
 * fp64 restatement, one world at a time (OpenMP across worlds), of
 *   /root/reference/src/environments/base_env.py      (step :354-413, termination :521-554,
 *                                                       reset :286-352, spawn :123-166)
 *   /root/reference/src/agents/entity.py              (action :126-134, ray sensor :159-220,
 *                                                       classification :222-241)
 *   /root/reference/src/agents/cop.py:49-75, thief.py:48-69           (rewards)
 *   /root/reference/src/environments/observation_spaces.py:67-131     (team-shared merge)
 * plus the Chipmunk2D 7.0.3 routines those lines reach through pymunk (named at each function).
 *
 * Deliberate, documented simplifications of the engine (DESIGN.md "Oracle"):
 *   - angular state is dropped: friction 0 and central contacts keep w == 0 exactly in exact
 *     arithmetic, so r x n terms vanish and nMass is 1 (wall) or 1/2 (agent pair);
 *   - arbiter order is fixed (agent-major wall contacts by hull id, then agent pairs) where
 *     Chipmunk's is BB-tree order (unspecified);
 *   - the BB-tree leaf test of a segment query is applied per shape with t_exit = 1 (no
 *     order-dependent pruning); dynamic leaves are treated as always visited.
 */
#include "cat_oracle.h"

#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/cat_philox.h"

#define MAX_AGENTS 8
#define TYPE_WALL 0
#define TYPE_COP 1
#define TYPE_THIEF 2
#define TYPE_EMPTY 4

typedef struct { double x, y; } V2;

static inline V2 v2(double x, double y) { V2 r = {x, y}; return r; }
static inline V2 vadd(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline V2 vsub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline V2 vmul(V2 a, double s) { return v2(a.x * s, a.y * s); }
static inline double vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline double vcross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline double vlen(V2 a) { return sqrt(vdot(a, a)); }
static inline V2 vlerp(V2 a, V2 b, double t) { return vadd(vmul(a, 1.0 - t), vmul(b, t)); }
static inline V2 vnormalize(V2 a) { return vmul(a, 1.0 / (vlen(a) + DBL_MIN)); } /* cpvnormalize */

struct OrcEnv {
  OrcParams p;
  int H, E, A, nc, nt, R;
  int* hull_off;
  V2* vert;    /* planes[i].v0 */
  V2* normal;  /* planes[i].n: outward normal of edge vert[i-1] -> vert[i] (cpPolyShape.c SetVerts) */
  double* bb;  /* [H][4] l,b,r,t of the shape = hull AABB grown by wall_radius (cpPolyShapeCacheData) */
  V2* init_pos;
  int* region_off;
  double* regions;
  double* ray_cos;
  double* ray_sin;
};

/* ------------------------------------------------------------------ half precision helpers */
uint16_t orc_double_to_half_bits(double x) {
  _Float16 h = (_Float16)x; /* single correctly-rounded conversion, like numpy's double->half */
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
static inline uint16_t float_to_half_bits(float x) {
  _Float16 h = (_Float16)x;
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
float orc_half_bits_to_float(uint16_t b) {
  _Float16 h;
  memcpy(&h, &b, 2);
  return (float)h;
}

void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out4) {
  cat_u32x4 r = cat_philox4x32_10(c0, c1, c2, c3, k0, k1);
  memcpy(out4, r.v, 16);
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ create / destroy */
OrcEnv* orc_create(const OrcMap* m, const OrcParams* p) {
  OrcEnv* e = (OrcEnv*)calloc(1, sizeof(OrcEnv));
  e->p = *p;
  e->H = m->n_hulls;
  e->E = m->n_edges;
  e->nc = m->n_cops;
  e->nt = m->n_thieves;
  e->A = e->nc + e->nt;
  e->R = p->n_rays;
  if (e->A > MAX_AGENTS || e->A < 1 || e->R > 512 || e->R < 1) { free(e); return NULL; }
  e->hull_off = (int*)malloc(sizeof(int) * (e->H + 1));
  memcpy(e->hull_off, m->hull_off, sizeof(int) * (e->H + 1));
  e->vert = (V2*)malloc(sizeof(V2) * e->E);
  e->normal = (V2*)malloc(sizeof(V2) * e->E);
  e->bb = (double*)malloc(sizeof(double) * 4 * e->H);
  for (int i = 0; i < e->E; ++i) e->vert[i] = v2(m->vert[2 * i], m->vert[2 * i + 1]);
  for (int h = 0; h < e->H; ++h) {
    int o = e->hull_off[h], n = e->hull_off[h + 1] - o;
    double l = INFINITY, b = INFINITY, r = -INFINITY, t = -INFINITY;
    for (int i = 0; i < n; ++i) {
      V2 a = e->vert[o + (i - 1 + n) % n], bb = e->vert[o + i];
      V2 d = vsub(bb, a);
      e->normal[o + i] = vnormalize(v2(d.y, -d.x)); /* cpvnormalize(cpvrperp(b - a)) */
      l = fmin(l, bb.x); r = fmax(r, bb.x); b = fmin(b, bb.y); t = fmax(t, bb.y);
    }
    double rad = p->wall_radius;
    e->bb[4 * h + 0] = l - rad; e->bb[4 * h + 1] = b - rad;
    e->bb[4 * h + 2] = r + rad; e->bb[4 * h + 3] = t + rad;
  }
  e->init_pos = (V2*)malloc(sizeof(V2) * e->A);
  for (int a = 0; a < e->A; ++a) e->init_pos[a] = v2(m->init_pos[2 * a], m->init_pos[2 * a + 1]);
  e->region_off = (int*)malloc(sizeof(int) * (e->A + 1));
  memcpy(e->region_off, m->region_off, sizeof(int) * (e->A + 1));
  int nr = e->region_off[e->A];
  e->regions = (double*)malloc(sizeof(double) * 4 * (nr > 0 ? nr : 1));
  if (nr > 0) memcpy(e->regions, m->regions, sizeof(double) * 4 * nr);
  /* entity.py:182-184: angles = linspace(0, 2*pi, n_rays, endpoint=False); cos, sin */
  e->ray_cos = (double*)malloc(sizeof(double) * e->R);
  e->ray_sin = (double*)malloc(sizeof(double) * e->R);
  double step = (2.0 * M_PI - 0.0) / (double)e->R; /* numpy linspace: start + i*step */
  for (int i = 0; i < e->R; ++i) {
    double ang = (double)i * step;
    e->ray_cos[i] = cos(ang);
    e->ray_sin[i] = sin(ang);
  }
  return e;
}

void orc_destroy(OrcEnv* e) {
  if (!e) return;
  free(e->hull_off); free(e->vert); free(e->normal); free(e->bb); free(e->init_pos);
  free(e->region_off); free(e->regions); free(e->ray_cos); free(e->ray_sin);
  free(e);
}

void orc_init_state(const OrcEnv* e, OrcState* st) {
  int A = e->A, H = e->H;
  for (int w = 0; w < st->n_worlds; ++w) {
    for (int a = 0; a < A; ++a) {
      size_t i = ((size_t)w * A + a) * 2;
      st->pos[i] = st->tc[i] = e->init_pos[a].x;         /* entity.py:115 body.position = start; */
      st->pos[i + 1] = st->tc[i + 1] = e->init_pos[a].y; /* space.add caches the shape centre (A.10) */
      st->vel[i] = st->vel[i + 1] = st->vbias[i] = st->vbias[i + 1] = 0.0;
    }
    st->step_count[w] = 0;
    st->episode[w] = 0;
    for (size_t k = 0; k < (size_t)A * H; ++k) { st->wall_jn[(size_t)w * A * H + k] = 0; st->wall_age[(size_t)w * A * H + k] = -1; }
    for (size_t k = 0; k < (size_t)A * A; ++k) { st->pair_jn[(size_t)w * A * A + k] = 0; st->pair_age[(size_t)w * A * A + k] = -1; }
  }
}

/* ------------------------------------------------------------------ Chipmunk geometry */

/* cpBBSegmentQuery (cpBB.h): entry fraction of segment a->b into bb, INFINITY on a miss. */
static double bb_segment_query(const double* bb, V2 a, V2 b) {
  V2 delta = vsub(b, a);
  double tmin = -INFINITY, tmax = INFINITY;
  if (delta.x == 0.0) {
    if (a.x < bb[0] || bb[2] < a.x) return INFINITY;
  } else {
    double t1 = (bb[0] - a.x) / delta.x, t2 = (bb[2] - a.x) / delta.x;
    tmin = fmax(tmin, fmin(t1, t2));
    tmax = fmin(tmax, fmax(t1, t2));
  }
  if (delta.y == 0.0) {
    if (a.y < bb[1] || bb[3] < a.y) return INFINITY;
  } else {
    double t1 = (bb[1] - a.y) / delta.y, t2 = (bb[3] - a.y) / delta.y;
    tmin = fmax(tmin, fmin(t1, t2));
    tmax = fmin(tmax, fmax(t1, t2));
  }
  if (tmin <= tmax && 0.0 <= tmax && tmin <= 1.0) return fmax(tmin, 0.0);
  return INFINITY;
}

/* cpClosetPointOnSegment (cpVect.h) */
static V2 closest_on_segment(V2 p, V2 a, V2 b) {
  V2 delta = vsub(a, b);
  double t = vdot(delta, vsub(p, b)) / vdot(delta, delta);
  t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
  return vadd(b, vmul(delta, t));
}

/* cpPolyShapePointQuery (cpPolyShape.c): signed distance to the RAW hull (caller subtracts r). */
static double hull_signed_distance(const OrcEnv* e, int h, V2 p) {
  int o = e->hull_off[h], n = e->hull_off[h + 1] - o;
  V2 v0 = e->vert[o + n - 1];
  double minDist = INFINITY;
  int outside = 0;
  for (int i = 0; i < n; ++i) {
    V2 v1 = e->vert[o + i];
    outside = outside || (vdot(e->normal[o + i], vsub(p, v1)) > 0.0);
    V2 c = closest_on_segment(p, v0, v1);
    double d = vlen(vsub(p, c));
    if (d < minDist) minDist = d;
    v0 = v1;
  }
  return outside ? minDist : -minDist;
}

double orc_hull_distance(const OrcEnv* e, int h, double px, double py) { return hull_signed_distance(e, h, v2(px, py)); }

/* CircleSegmentQuery (chipmunk_private.h) */
static int circle_segment_query(V2 center, double r1, V2 a, V2 b, double r2, double* alpha, V2* point) {
  V2 da = vsub(a, center), db = vsub(b, center);
  double rsum = r1 + r2;
  double qa = vdot(da, da) - 2.0 * vdot(da, db) + vdot(db, db);
  double qb = vdot(da, db) - vdot(da, da);
  double det = qb * qb - qa * (vdot(da, da) - rsum * rsum);
  if (det >= 0.0) {
    double t = (-qb - sqrt(det)) / qa;
    if (0.0 <= t && t <= 1.0) {
      V2 n = vnormalize(vlerp(da, db, t));
      *alpha = t;
      *point = vsub(vlerp(a, b, t), vmul(n, r2));
      return 1;
    }
  }
  return 0;
}

/* cpShapeSegmentQuery (cpShape.c) + cpPolyShapeSegmentQuery (cpPolyShape.c) for hull h. */
static int poly_shape_segment_query(const OrcEnv* e, int h, V2 a, V2 b, double r2, double* alpha, V2* point) {
  double r = e->p.wall_radius;
  /* cpShapeSegmentQuery: start point within `radius` of the shape -> alpha 0, point stays at b */
  double nearest = hull_signed_distance(e, h, a) - r;
  if (nearest <= r2) { *alpha = 0.0; *point = b; return 1; }

  int o = e->hull_off[h], n = e->hull_off[h + 1] - o;
  double rsum = r + r2;
  int hit = 0;
  double best = 1.0;
  V2 bestp = b;
  for (int i = 0; i < n; ++i) {
    V2 nn = e->normal[o + i];
    double an = vdot(a, nn);
    double d = an - vdot(e->vert[o + i], nn) - rsum;
    if (d < 0.0) continue;
    double bn = vdot(b, nn);
    double t = d / fmax(an - bn, DBL_MIN);
    if (t < 0.0 || 1.0 < t) continue;
    V2 pt = vlerp(a, b, t);
    double dt = vcross(nn, pt);
    double dtMin = vcross(nn, e->vert[o + (i - 1 + n) % n]);
    double dtMax = vcross(nn, e->vert[o + i]);
    if (dtMin <= dt && dt <= dtMax) {
      hit = 1; best = t; bestp = vsub(vlerp(a, b, t), vmul(nn, r2)); /* last passing plane wins */
    }
  }
  if (rsum > 0.0) { /* bevelled vertexes */
    for (int i = 0; i < n; ++i) {
      double ca; V2 cp;
      if (circle_segment_query(e->vert[o + i], r, a, b, r2, &ca, &cp) && ca < best) { hit = 1; best = ca; bestp = cp; }
    }
  }
  if (hit) { *alpha = best; *point = bestp; }
  return hit;
}

/* cpShapeSegmentQuery for an agent circle whose cached world centre is c. */
static int circle_shape_segment_query(V2 c, double r, V2 a, V2 b, double r2, double* alpha, V2* point) {
  double nearest = vlen(vsub(a, c)) - r; /* cpCircleShapePointQuery */
  if (nearest <= r2) { *alpha = 0.0; *point = b; return 1; }
  return circle_segment_query(c, r, a, b, r2, alpha, point);
}

/* cpSpaceSegmentQueryFirst (cpSpaceQuery.c): static index first, then dynamic, strict '<'.
 * self_agent >= 0: sensor filter (own shape rejected by group, entity.py:120-123);
 * self_agent  < 0: capture LOS filter (both agent categories masked out, base_env.py:536-538). */
static int segment_query_first(const OrcEnv* e, const double* tc, int self_agent, V2 a, V2 b, double radius,
                               double* out_alpha, V2* out_point) {
  double best = 1.0;
  int shape = -1;
  V2 bp = b;
  for (int h = 0; h < e->H; ++h) {
    /* cpBBTree SubtreeSegmentQuery: a leaf is visited only if the THIN segment enters its bb
     * (the spatial index knows nothing about the query radius) with entry fraction < t_exit. */
    if (!(bb_segment_query(&e->bb[4 * h], a, b) < 1.0)) continue;
    double al; V2 pt;
    if (poly_shape_segment_query(e, h, a, b, radius, &al, &pt) && al < best) { best = al; shape = h; bp = pt; }
  }
  if (self_agent >= 0) {
    for (int j = 0; j < e->A; ++j) {
      if (j == self_agent) continue;
      double al; V2 pt;
      if (circle_shape_segment_query(v2(tc[2 * j], tc[2 * j + 1]), e->p.unit_size, a, b, radius, &al, &pt) && al < best) {
        best = al; shape = e->H + j; bp = pt;
      }
    }
  }
  *out_alpha = best;
  *out_point = bp;
  return shape;
}

int orc_segment_query_first(const OrcEnv* env, const double* tc, int self_agent, double ax, double ay, double bx,
                            double by, double radius, double* alpha, double* point) {
  V2 p;
  int s = segment_query_first(env, tc, self_agent, v2(ax, ay), v2(bx, by), radius, alpha, &p);
  point[0] = p.x; point[1] = p.y;
  return s;
}

/* ------------------------------------------------------------------ observation */
typedef struct {
  uint16_t dist[MAX_AGENTS][512];
  uint8_t type[MAX_AGENTS][512];
} ObsBuf;

/* entity.py:159-220 for agent a of one world */
static void agent_observation(const OrcEnv* e, const double* pos, const double* tc, int a, uint16_t* dist,
                              uint8_t* type, double* hit_point, double* hit_alpha) {
  V2 origin = v2(pos[2 * a], pos[2 * a + 1]);
  double L = e->p.ray_length;
  uint16_t ox16 = orc_double_to_half_bits(origin.x), oy16 = orc_double_to_half_bits(origin.y);
  float oxf = orc_half_bits_to_float(ox16), oyf = orc_half_bits_to_float(oy16);
  for (int i = 0; i < e->R; ++i) {
    V2 end = v2(origin.x + L * e->ray_cos[i], origin.y + L * e->ray_sin[i]); /* entity.py:191-193 */
    double alpha; V2 pt;
    int shape = segment_query_first(e, tc, a, origin, end, e->p.ray_radius, &alpha, &pt);
    if (hit_point) { hit_point[2 * i] = pt.x; hit_point[2 * i + 1] = pt.y; }
    if (hit_alpha) hit_alpha[i] = alpha;
    if (shape < 0) { /* entity.py:200-201 */
      dist[i] = orc_double_to_half_bits(L);
      type[i] = TYPE_EMPTY;
      continue;
    }
    /* entity.py:206-210: points -> f16; dx,dy in f16 (python-float origin is cast to f16); hypot in f16 */
    float pxf = orc_half_bits_to_float(orc_double_to_half_bits(pt.x));
    float pyf = orc_half_bits_to_float(orc_double_to_half_bits(pt.y));
    float dx = orc_half_bits_to_float(float_to_half_bits(pxf - oxf)); /* numpy half subtract: float op, round to half */
    float dy = orc_half_bits_to_float(float_to_half_bits(pyf - oyf));
    float hyp = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy); /* npy_hypotf, then -> half */
    dist[i] = float_to_half_bits(hyp);
    /* entity.py:222-241 */
    if (shape < e->H) type[i] = TYPE_WALL;
    else type[i] = ((shape - e->H) >= e->nc) ? TYPE_THIEF : TYPE_COP;
  }
}

/* cop.py:49-75 / thief.py:48-69, evaluated in double from the f16 distances. */
static float agent_reward(const OrcEnv* e, int a, const uint16_t* dist, const uint8_t* type, int captured, int timeout) {
  int is_cop = a < e->nc;
  if (captured) return is_cop ? 1.0f : -1.0f;
  if (timeout) return is_cop ? -1.0f : 1.0f;
  int want = is_cop ? TYPE_THIEF : TYPE_COP;
  int seen = 0;
  float dmin = INFINITY;
  for (int i = 0; i < e->R; ++i)
    if (type[i] == want) { float d = orc_half_bits_to_float(dist[i]); if (!seen || d < dmin) dmin = d; seen = 1; }
  if (is_cop) {
    double r = -0.02;
    if (seen) r += 1.5 * exp(-(double)dmin / 50.0); else r -= 0.02;
    return (float)r;
  }
  if (seen) return (float)(tanh(((double)dmin - 100.0) / 50.0) / 10.0);
  return 0.15f;
}

/* observation_spaces.py:67-131 net effect: per ray first non-EMPTY (type, distance) in team order. */
static void shared_observation(const OrcEnv* e, const ObsBuf* ob, const double* pos, uint16_t* shared_dist,
                               uint8_t* shared_type, uint16_t* team_pos) {
  int R = e->R;
  for (int team = 0; team < 2; ++team) {
    int a0 = team == 0 ? 0 : e->nc, a1 = team == 0 ? e->nc : e->A;
    for (int i = 0; i < R; ++i) {
      uint8_t t = TYPE_EMPTY;
      uint16_t d = 0; /* dist_masked = zeros_like(...) — always overwritten when the team is non-empty */
      for (int a = a0; a < a1; ++a) {
        if (t == TYPE_EMPTY) { t = ob->type[a][i]; d = ob->dist[a][i]; }
      }
      if (shared_type) shared_type[team * R + i] = t;
      if (shared_dist) shared_dist[team * R + i] = d;
    }
  }
  if (team_pos)
    for (int a = 0; a < e->A; ++a) { /* observation_spaces.py:92-95: float16(body.position) */
      team_pos[2 * a] = orc_double_to_half_bits(pos[2 * a]);
      team_pos[2 * a + 1] = orc_double_to_half_bits(pos[2 * a + 1]);
    }
}

static void observe_world(const OrcEnv* e, const double* pos, const double* tc, int w, OrcOut* out, ObsBuf* ob) {
  int A = e->A, R = e->R;
  for (int a = 0; a < A; ++a) {
    size_t o = ((size_t)w * A + a) * R;
    agent_observation(e, pos, tc, a, ob->dist[a], ob->type[a], out->hit_point ? out->hit_point + 2 * o : NULL,
                      out->hit_alpha ? out->hit_alpha + o : NULL);
    if (out->obs_dist) memcpy(out->obs_dist + o, ob->dist[a], sizeof(uint16_t) * R);
    if (out->obs_type) memcpy(out->obs_type + o, ob->type[a], R);
  }
  shared_observation(e, ob, pos, out->shared_dist ? out->shared_dist + (size_t)w * 2 * R : NULL,
                     out->shared_type ? out->shared_type + (size_t)w * 2 * R : NULL,
                     out->team_pos ? out->team_pos + (size_t)w * A * 2 : NULL);
}

void orc_observe(const OrcEnv* e, const OrcState* st, OrcOut* out) {
  int A = e->A;
#pragma omp parallel for schedule(static)
  for (int w = 0; w < st->n_worlds; ++w) {
    ObsBuf ob;
    observe_world(e, st->pos + (size_t)w * A * 2, st->tc + (size_t)w * A * 2, w, out, &ob);
  }
}

/* ------------------------------------------------------------------ termination: base_env.py:521-554 */
static void termination_criterion(const OrcEnv* e, const double* pos, const double* tc, int step_count, int* captured,
                                  int* timeout) {
  *captured = 0; *timeout = 0;
  for (int t = e->nc; t < e->A; ++t)
    for (int c = 0; c < e->nc; ++c) {
      V2 tp = v2(pos[2 * t], pos[2 * t + 1]), cp = v2(pos[2 * c], pos[2 * c + 1]);
      double al; V2 pt;
      int hit = segment_query_first(e, tc, -1, tp, cp, 0.0, &al, &pt);
      if (hit < 0 && vlen(vsub(tp, cp)) < e->p.termination_radius) { *captured = 1; return; }
    }
  if (step_count >= e->p.max_step_count) *timeout = 1;
}

/* ------------------------------------------------------------------ physics: cpSpaceStep */
typedef struct { int a, b; /* b = -1: static wall */ V2 n; double dist, nMass, bias, jBias, jnAcc; int first; double* cache_jn; int8_t* cache_age; } Contact;

/* Closest points between an agent centre (a point; CircleToPoly's GJK/EPA support shape) and the raw
 * hull (cpCollision.c CircleToPoly + ClosestPointsNew): n points from the circle towards the hull,
 * d is the signed centre-to-hull distance (negative inside: least-penetration edge, EPA). */
static void circle_hull_closest(const OrcEnv* e, int h, V2 c, V2* n_out, double* d_out) {
  int o = e->hull_off[h], n = e->hull_off[h + 1] - o;
  int inside = 1, imax = 0, best_edge = 0, best_interior = 0;
  double maxpd = -INFINITY, bestd2 = INFINITY;
  V2 bestq = c;
  for (int i = 0; i < n; ++i) {
    V2 v1 = e->vert[o + i], v0 = e->vert[o + (i - 1 + n) % n];
    double pd = vdot(e->normal[o + i], vsub(c, v1));
    if (pd > 0.0) inside = 0;
    if (pd > maxpd) { maxpd = pd; imax = i; }
    V2 ed = vsub(v1, v0);
    double t = vdot(vsub(c, v0), ed) / vdot(ed, ed);
    int interior = (t > 0.0 && t < 1.0);
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    V2 q = vadd(v0, vmul(ed, t));
    double d2 = vdot(vsub(c, q), vsub(c, q));
    if (d2 < bestd2) { bestd2 = d2; bestq = q; best_edge = i; best_interior = interior; }
  }
  if (inside) { *d_out = maxpd; *n_out = vmul(e->normal[o + imax], -1.0); return; }
  double d = sqrt(bestd2);
  *d_out = d;
  if (best_interior) *n_out = vmul(e->normal[o + best_edge], -1.0); /* edge region: MSA = edge normal */
  else *n_out = vmul(vsub(bestq, c), 1.0 / (d + DBL_MIN));          /* vertex region: p/|p| */
}

static void physics_step(const OrcEnv* e, double* pos, double* vel, double* vbias, double* tc, double* wall_jn,
                         int8_t* wall_age, double* pair_jn, int8_t* pair_age) {
  const OrcParams* p = &e->p;
  int A = e->A, H = e->H;
  double dt = p->dt;
  /* (1) cpBodyUpdatePosition: p += (v + v_bias)*dt; v_bias = 0.  (2) shape caches refreshed. */
  for (int a = 0; a < A; ++a) {
    pos[2 * a] += (vel[2 * a] + vbias[2 * a]) * dt;
    pos[2 * a + 1] += (vel[2 * a + 1] + vbias[2 * a + 1]) * dt;
    vbias[2 * a] = vbias[2 * a + 1] = 0.0;
    tc[2 * a] = pos[2 * a]; tc[2 * a + 1] = pos[2 * a + 1];
  }
  /* (3) narrow phase -> arbiters.  Order: wall contacts agent-major by hull id, then agent pairs. */
  Contact* cons = (Contact*)malloc(sizeof(Contact) * ((size_t)A * H + (size_t)A * A));
  int nc = 0;
  double rsum_w = p->unit_size + p->wall_radius;
  for (int a = 0; a < A; ++a) {
    V2 c = v2(pos[2 * a], pos[2 * a + 1]);
    for (int h = 0; h < H; ++h) {
      /* QueryReject: shape BBs must overlap (inclusive) — implied by d <= rsum, kept for fidelity */
      const double* bb = &e->bb[4 * h];
      double r = p->unit_size;
      if (!(c.x - r <= bb[2] && bb[0] <= c.x + r && c.y - r <= bb[3] && bb[1] <= c.y + r)) continue;
      V2 n; double d;
      circle_hull_closest(e, h, c, &n, &d);
      if (d <= rsum_w) { /* CircleToPoly: points.d <= circle->r + poly->r */
        Contact* k = &cons[nc++];
        k->a = a; k->b = -1; k->n = n; k->dist = d - rsum_w; k->nMass = 1.0 / (1.0 / p->unit_mass);
        k->cache_jn = &wall_jn[(size_t)a * H + h]; k->cache_age = &wall_age[(size_t)a * H + h];
      }
    }
  }
  for (int i = 0; i < A; ++i)
    for (int j = i + 1; j < A; ++j) {
      /* CircleToCircle: strict distsq < mindist^2; n = (1,0) if coincident */
      V2 delta = v2(pos[2 * j] - pos[2 * i], pos[2 * j + 1] - pos[2 * i + 1]);
      double mind = 2.0 * p->unit_size, dsq = vdot(delta, delta);
      if (dsq < mind * mind) {
        double dist = sqrt(dsq);
        Contact* k = &cons[nc++];
        k->a = i; k->b = j; k->n = dist ? vmul(delta, 1.0 / dist) : v2(1.0, 0.0);
        k->dist = dist - mind; k->nMass = 1.0 / (2.0 / p->unit_mass);
        k->cache_jn = &pair_jn[(size_t)i * A + j]; k->cache_age = &pair_age[(size_t)i * A + j];
      }
    }
  /* cpArbiterUpdate: contact hash of circle contacts is always 0 -> jnAcc persists while the
   * arbiter is cached; state FIRST_COLLISION unless it was used in the previous step. */
  for (int k = 0; k < nc; ++k) {
    Contact* c = &cons[k];
    int age = *c->cache_age;
    c->jnAcc = age >= 0 ? *c->cache_jn : 0.0;
    c->first = (age != 0);
    *c->cache_age = -2; /* mark "used this step" */
  }
  /* (4) cpSpaceArbiterSetFilter: unused arbiters age; dropped once ticks >= collision_persistence */
  for (size_t k = 0; k < (size_t)A * H; ++k) {
    if (wall_age[k] == -2) wall_age[k] = 0;
    else if (wall_age[k] >= 0) { wall_age[k]++; if (wall_age[k] >= p->collision_persistence) { wall_age[k] = -1; wall_jn[k] = 0.0; } }
  }
  for (size_t k = 0; k < (size_t)A * A; ++k) {
    if (pair_age[k] == -2) pair_age[k] = 0;
    else if (pair_age[k] >= 0) { pair_age[k]++; if (pair_age[k] >= p->collision_persistence) { pair_age[k] = -1; pair_jn[k] = 0.0; } }
  }
  /* (5) cpArbiterPreStep */
  double biasCoef = 1.0 - pow(p->collision_bias, dt);
  for (int k = 0; k < nc; ++k) {
    Contact* c = &cons[k];
    c->bias = -biasCoef * fmin(0.0, c->dist + p->collision_slop) / dt;
    c->jBias = 0.0;
  }
  /* (6) cpBodyUpdateVelocity: damping 1, gravity 0, no forces -> identity */
  /* (7) cpArbiterApplyCachedImpulse, dt_coef = dt/prev_dt = 1 (constant dt; nothing is cached on
   *     the first step of a space, where prev_dt = 0) */
  double minv = 1.0 / p->unit_mass;
  for (int k = 0; k < nc; ++k) {
    Contact* c = &cons[k];
    if (c->first) continue;
    vel[2 * c->a] -= c->n.x * c->jnAcc * minv; vel[2 * c->a + 1] -= c->n.y * c->jnAcc * minv;
    if (c->b >= 0) { vel[2 * c->b] += c->n.x * c->jnAcc * minv; vel[2 * c->b + 1] += c->n.y * c->jnAcc * minv; }
  }
  /* (8) cpArbiterApplyImpulse x iterations (e = 0, u = 0: normal impulses only) */
  for (int it = 0; it < p->iterations; ++it)
    for (int k = 0; k < nc; ++k) {
      Contact* c = &cons[k];
      V2 n = c->n;
      V2 va = v2(vel[2 * c->a], vel[2 * c->a + 1]), vba = v2(vbias[2 * c->a], vbias[2 * c->a + 1]);
      V2 vb = v2(0, 0), vbb = v2(0, 0);
      if (c->b >= 0) { vb = v2(vel[2 * c->b], vel[2 * c->b + 1]); vbb = v2(vbias[2 * c->b], vbias[2 * c->b + 1]); }
      double vbn = vdot(vsub(vbb, vba), n);
      double vrn = vdot(vsub(vb, va), n);
      double jbn = (c->bias - vbn) * c->nMass;
      double jbnOld = c->jBias;
      c->jBias = fmax(jbnOld + jbn, 0.0);
      double jn = -(0.0 + vrn) * c->nMass;
      double jnOld = c->jnAcc;
      c->jnAcc = fmax(jnOld + jn, 0.0);
      double jb = c->jBias - jbnOld, j = c->jnAcc - jnOld;
      vbias[2 * c->a] -= n.x * jb * minv; vbias[2 * c->a + 1] -= n.y * jb * minv;
      vel[2 * c->a] -= n.x * j * minv; vel[2 * c->a + 1] -= n.y * j * minv;
      if (c->b >= 0) {
        vbias[2 * c->b] += n.x * jb * minv; vbias[2 * c->b + 1] += n.y * jb * minv;
        vel[2 * c->b] += n.x * j * minv; vel[2 * c->b + 1] += n.y * j * minv;
      }
    }
  for (int k = 0; k < nc; ++k) *cons[k].cache_jn = cons[k].jnAcc;
  free(cons);
}

/* ------------------------------------------------------------------ reset: base_env.py:286-352 */
/* _get_non_colliding_position (base_env.py:123-166) with Philox in place of np_random/random. */
static V2 sample_spawn(const OrcEnv* e, const double* tc, int a, uint64_t gid, uint32_t episode) {
  int r0 = e->region_off[a], nr = e->region_off[a + 1] - r0;
  uint32_t idx = cat_spawn_region_index(e->p.seed, gid, episode, (uint32_t)a, (uint32_t)nr);
  const double* reg = &e->regions[4 * (r0 + (int)idx)];
  float rx = (float)reg[0], ry = (float)reg[1], rw = (float)reg[2], rh = (float)reg[3];
  for (uint32_t t = 0; t < 20; ++t) {
    float ux, uy;
    cat_spawn_uniforms(e->p.seed, gid, episode, (uint32_t)a, t, &ux, &uy);
    /* fp32 fma so the GPU reproduces the sampled point bit-for-bit */
    V2 pos = v2((double)fmaf(rw, ux, rx), (double)fmaf(rh, uy, ry));
    /* space.point_query_nearest(pos, radius 5, ray_filter): any shape with distance < 5 */
    int blocked = 0;
    for (int h = 0; h < e->H && !blocked; ++h)
      if (hull_signed_distance(e, h, pos) - e->p.wall_radius < e->p.unit_size) blocked = 1;
    for (int j = 0; j < e->A && !blocked; ++j) {
      if (j == a) continue; /* own shape rejected by group */
      if (vlen(vsub(pos, v2(tc[2 * j], tc[2 * j + 1]))) - e->p.unit_size < e->p.unit_size) blocked = 1;
    }
    if (!blocked) return pos;
  }
  return v2((double)(rx + rw / 2.0f), (double)(ry + rh / 2.0f)); /* base_env.py:163-166 */
}

static void reset_world(const OrcEnv* e, OrcState* st, int w) {
  int A = e->A;
  double* pos = st->pos + (size_t)w * A * 2;
  double* vel = st->vel + (size_t)w * A * 2;
  double* tc = st->tc + (size_t)w * A * 2;
  st->episode[w] += 1;
  for (int a = 0; a < A; ++a) {
    V2 np_;
    if (e->region_off[a + 1] > e->region_off[a]) np_ = sample_spawn(e, tc, a, (uint64_t)(st->gid0 + w), st->episode[w]);
    else np_ = e->init_pos[a]; /* base_env.py:328-332 -> Entity.reset() to its initial position */
    pos[2 * a] = np_.x; pos[2 * a + 1] = np_.y; /* entity.py:154-156 */
    vel[2 * a] = vel[2 * a + 1] = 0.0;          /* entity.py:157 */
    /* v_bias, cached arbiters and (pymunk, A.10) the cached shape centre are NOT touched */
    if (!e->p.stale_shape_cache) { tc[2 * a] = np_.x; tc[2 * a + 1] = np_.y; }
  }
  st->step_count[w] = 0; /* base_env.py:350 */
}

void orc_reset(const OrcEnv* e, OrcState* st, const uint8_t* mask, OrcOut* out) {
  int A = e->A;
#pragma omp parallel for schedule(static)
  for (int w = 0; w < st->n_worlds; ++w) {
    if (mask && !mask[w]) continue;
    reset_world(e, st, w);
    if (out) { ObsBuf ob; observe_world(e, st->pos + (size_t)w * A * 2, st->tc + (size_t)w * A * 2, w, out, &ob); }
  }
}

/* ------------------------------------------------------------------ step: base_env.py:354-413 */
void orc_step(const OrcEnv* e, OrcState* st, const int32_t* actions, OrcOut* out) {
  int A = e->A, H = e->H, R = e->R;
  const OrcParams* p = &e->p;
#pragma omp parallel for schedule(static)
  for (int w = 0; w < st->n_worlds; ++w) {
    double* pos = st->pos + (size_t)w * A * 2;
    double* vel = st->vel + (size_t)w * A * 2;
    double* vbias = st->vbias + (size_t)w * A * 2;
    double* tc = st->tc + (size_t)w * A * 2;
    ObsBuf ob;
    st->step_count[w] += 1; /* :372 */
    int captured, timeout;
    termination_criterion(e, pos, tc, st->step_count[w], &captured, &timeout); /* :378 */
    for (int a = 0; a < A; ++a) { /* :380-383 -> Entity.step -> _perform_action (entity.py:126-134) */
      int act = actions[(size_t)w * A + a];
      double fx = 0, fy = 0, s = p->unit_velocity;
      if (act == 0) fx = -s; else if (act == 1) fy = s; else if (act == 2) fx = s; else if (act == 3) fy = -s;
      double vx = vel[2 * a] + fx / p->unit_mass, vy = vel[2 * a + 1] + fy / p->unit_mass;
      double sp = sqrt(vx * vx + vy * vy);
      if (sp > p->max_speed) { vx = vx / sp * p->max_speed; vy = vy / sp * p->max_speed; }
      vel[2 * a] = vx; vel[2 * a + 1] = vy;
    }
    observe_world(e, pos, tc, w, out, &ob); /* entity.py:143 + base_env.py:388 */
    for (int a = 0; a < A; ++a)
      if (out->reward) out->reward[(size_t)w * A + a] = agent_reward(e, a, ob.dist[a], ob.type[a], captured, timeout);
    physics_step(e, pos, vel, vbias, tc, st->wall_jn + (size_t)w * A * H, st->wall_age + (size_t)w * A * H,
                 st->pair_jn + (size_t)w * A * A, st->pair_age + (size_t)w * A * A); /* :392 */
    int done = captured || timeout;
    if (out->terminated) out->terminated[w] = (uint8_t)done;
    if (out->truncated) out->truncated[w] = (uint8_t)timeout;
    if (out->winner) out->winner[w] = done ? (captured ? 0 : 1) : -1;
    if (done && p->auto_reset) { /* the trainer's reset(), obs of the reset state */
      reset_world(e, st, w);
      observe_world(e, pos, tc, w, out, &ob);
    }
    (void)R;
  }
}

/* space.step(dt) alone (base_env.py:392 -> cpSpaceStep), every world: what the stand-in pymunk of the harness
 * self-test (tests/fake_pymunk) and the multi-step checks call. */
void orc_space_step(const OrcEnv* e, OrcState* st) {
  int A = e->A, H = e->H;
#pragma omp parallel for schedule(static)
  for (int w = 0; w < st->n_worlds; ++w)
    physics_step(e, st->pos + (size_t)w * A * 2, st->vel + (size_t)w * A * 2, st->vbias + (size_t)w * A * 2,
                 st->tc + (size_t)w * A * 2, st->wall_jn + (size_t)w * A * H, st->wall_age + (size_t)w * A * H,
                 st->pair_jn + (size_t)w * A * A, st->pair_age + (size_t)w * A * A);
}

/* space.point_query_nearest(p, max_distance, filter) (cpSpacePointQueryNearest): nearest shape with distance <
 * max_distance; walls = hull distance - wall radius, agents = centre distance - radius, own shape excluded.
 * Returns the shape id (-1 none; < H hull; H + j agent j) and its distance. */
int orc_point_query_nearest(const OrcEnv* e, const double* tc, int self_agent, double px, double py, double maxd,
                            double* dist_out) {
  int best = -1;
  double bd = maxd;
  for (int h = 0; h < e->H; ++h) {
    double d = hull_signed_distance(e, h, v2(px, py)) - e->p.wall_radius;
    if (d < bd) { bd = d; best = h; }
  }
  for (int j = 0; j < e->A; ++j) {
    if (j == self_agent) continue;
    double d = vlen(vsub(v2(px, py), v2(tc[2 * j], tc[2 * j + 1]))) - e->p.unit_size;
    if (d < bd) { bd = d; best = e->H + j; }
  }
  if (dist_out) *dist_out = bd;
  return best;
}

/* ------------------------------------------------------------------ GAE (skrl MAPPO._update) */
void orc_gae(const float* rewards, const uint8_t* dones, const float* values, const float* last_values, float* returns,
             float* advantages, int T, int M, double gamma, double lam, int normalize) {
  double* adv = (double*)malloc(sizeof(double) * (size_t)T * M);
  for (int m = 0; m < M; ++m) {
    double a = 0.0;
    for (int t = T - 1; t >= 0; --t) {
      double nv = (t == T - 1) ? (double)last_values[m] : (double)values[(size_t)(t + 1) * M + m];
      double nd = dones[(size_t)t * M + m] ? 0.0 : 1.0;
      a = (double)rewards[(size_t)t * M + m] - (double)values[(size_t)t * M + m] + gamma * nd * (nv + lam * a);
      adv[(size_t)t * M + m] = a;
    }
  }
  size_t n = (size_t)T * M;
  for (size_t i = 0; i < n; ++i) returns[i] = (float)(adv[i] + (double)values[i]);
  if (normalize) {
    double mean = 0; for (size_t i = 0; i < n; ++i) mean += adv[i]; mean /= (double)n;
    double var = 0; for (size_t i = 0; i < n; ++i) var += (adv[i] - mean) * (adv[i] - mean);
    double sd = n > 1 ? sqrt(var / (double)(n - 1)) : 0.0; /* torch.std: unbiased */
    for (size_t i = 0; i < n; ++i) advantages[i] = (float)((adv[i] - mean) / (sd + 1e-8));
  } else {
    for (size_t i = 0; i < n; ++i) advantages[i] = (float)adv[i];
  }
  free(adv);
}
