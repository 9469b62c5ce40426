/*
 * rthx_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP) of RayTraceHeatTransfer.jl's `method=:exchange` tracer, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg as the checker and the
 * reported CPU baseline.  The product path (raytraceheattransfer.jl_b200/csrc) never links or calls it.
 *
 * It follows the reference function by function and keeps the reference's point-location algorithm
 * (uniform grid + crossing-number point-in-polygon, bbox-prefilter fallback), nudges and branch order:
 *   computeExchangeFactorsBin   RT2D/ExchangeFactors2D/parallelRayTracing.jl:64-159
 *   emitSurfaceRay2D            RT2D/Shared2D/emitSurfaceRay2D.jl:1-27
 *   lambertSample2D             RT2D/Shared2D/lambertSample2D.jl:1-11
 *   emitVolumeRay2D             RT2D/Shared2D/emitVolumeRay2D.jl:1-33
 *   traceRay / Uniform / Variable   RT2D/Shared2D/traceRay.jl:1-147
 *   distToSurface2D             RT2D/Shared2D/distToSurface2D.jl:2-18
 *   findFace2D family           RT2D/Shared2D/findFace2D.jl:1-101
 *   getGlobalIndex2D            RT2D/Shared2D/getGlobalIndex2D.jl:1-14
 *   buildUniformGrid / buildOptimizedSpatialStructure   src/Domains/domains/spatialAccelerations.jl:2-89
 *   calculateInwardNormal       src/Domains/domains/calculateInwardNormal.jl:1-12
 * (RT2D = src/RayTracing/RayTracing2D).
 *
 * The ONE deliberate difference: the reference draws from Julia's unseeded task-local `rand()`; the oracle
 * draws the same variates, in the same order, from the counter-based Philox4x32-10 stream defined in
 * include/rthx.h so that it consumes exactly the numbers the CUDA kernels consume.
 *
 * Parity pin: Julia is not installed in this image, so the reference cannot be executed; the oracle is
 * pinned against the reference's own golden vectors for this path — the Crosbie & Schrenker (1984) centre-line
 * table of test/test_2d_grey.jl:25-33 with its rtol=0.05 norm test (:216), the circle-centre temperature of
 * test/test_triangle_mesh.jl:66-69, the parallel-plate flux of test/test_2d_grey_reflecting.jl:96-136 and the
 * diffusion-limit source function of test/test_2d_diffusion.jl:19-23,57-76 — and against closed-form known
 * answers (tests/test_oracle_*.py).  No golden F matrix or seed exists in the reference (its RNG is unseeded):
 * F itself is pinned statistically, through these.
 */
#include "rthx_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11; the Random123 reference constants)                          */
/* ------------------------------------------------------------------------------------------------ */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void rthx_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* rthx.h RNG contract: 52-bit uniform in (0,1) from two words, 23-bit float uniform in (0,1) from one. */
double rthx_oracle_u52(uint32_t lo, uint32_t hi) {
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return ((double)(x >> 12) + 0.5) * 0x1p-52;
}
float rthx_oracle_u23(uint32_t w) { return ((float)(w >> 9) + 0.5f) * 0x1p-23f; }
double rthx_oracle_u32(uint32_t w) { return ((double)w + 0.5) * 0x1p-32; }

typedef struct {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t w[8];
} draws_t;

/* ---- "reference-faithful" timing mode (bench.py's CPU baseline) -------------------------------
 * The reference draws every variate from Julia's task-local xoshiro256++ (`rand()`, traceRay.jl:25,79 etc.) and tallies a
 * row in a Dict{Int,Int} (parallelRayTracing.jl:104,124).  With the switch on, the oracle does the same: the words that feed
 * the samplers come from a per-thread xoshiro256++ (seeded per thread by splitmix64) instead of Philox, and each emitter row
 * is tallied in an open-addressing hash map that is flushed into the row when the emitter is done (:144-146).  Same
 * geometry code, same draw order; results are statistically equivalent but NOT the Philox contract — for timing only. */
static int g_faithful = 0;
void rthx_oracle_set_faithful(int on) { g_faithful = on ? 1 : 0; }
int rthx_oracle_get_faithful(void) { return g_faithful; }

static _Thread_local uint64_t t_xo[4];
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xoshiro_next(void) {                /* xoshiro256++ 1.0 (Blackman & Vigna), Julia's default RNG */
  uint64_t* s = t_xo;
  const uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl64(s[3], 45);
  return result;
}
static void xoshiro_seed(uint64_t seed, uint64_t stream) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (stream + 1);
  for (int i = 0; i < 4; ++i) {                             /* splitmix64 */
    z += 0x9E3779B97F4A7C15ull;
    uint64_t x = z;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    t_xo[i] = x ^ (x >> 31);
  }
}
static inline void xoshiro_words(uint32_t* w, int n64) {
  for (int i = 0; i < n64; ++i) { const uint64_t x = xoshiro_next(); w[2 * i] = (uint32_t)x; w[2 * i + 1] = (uint32_t)(x >> 32); }
}

/* Dict{Int,Int}-like row tally: open addressing, linear probing, grows at load factor 2/3 (like Julia's Dict) */
typedef struct { int32_t* keys; uint64_t* vals; uint32_t cap, n; } rowmap_t;
static void rowmap_init(rowmap_t* m) {
  m->cap = 16; m->n = 0;
  m->keys = (int32_t*)malloc(sizeof(int32_t) * m->cap); m->vals = (uint64_t*)malloc(sizeof(uint64_t) * m->cap);
  for (uint32_t i = 0; i < m->cap; ++i) m->keys[i] = -1;
}
static void rowmap_clear(rowmap_t* m) { for (uint32_t i = 0; i < m->cap; ++i) m->keys[i] = -1; m->n = 0; }
static void rowmap_free(rowmap_t* m) { free(m->keys); free(m->vals); }
static inline uint32_t rowmap_slot(const rowmap_t* m, int32_t key) {
  uint32_t i = ((uint32_t)key * 0x9E3779B1u) & (m->cap - 1);
  while (m->keys[i] != -1 && m->keys[i] != key) i = (i + 1) & (m->cap - 1);
  return i;
}
static void rowmap_add(rowmap_t* m, int32_t key) {
  uint32_t i = rowmap_slot(m, key);
  if (m->keys[i] == key) { m->vals[i]++; return; }
  if (3 * (m->n + 1) > 2 * m->cap) {
    rowmap_t g; g.cap = m->cap * 4; g.n = m->n;
    g.keys = (int32_t*)malloc(sizeof(int32_t) * g.cap); g.vals = (uint64_t*)malloc(sizeof(uint64_t) * g.cap);
    for (uint32_t k = 0; k < g.cap; ++k) g.keys[k] = -1;
    for (uint32_t k = 0; k < m->cap; ++k)
      if (m->keys[k] != -1) { const uint32_t j = rowmap_slot(&g, m->keys[k]); g.keys[j] = m->keys[k]; g.vals[j] = m->vals[k]; }
    rowmap_free(m);
    *m = g;
    i = rowmap_slot(m, key);
  }
  m->keys[i] = key; m->vals[i] = 1; m->n++;
}

static void draws_init(draws_t* d, uint64_t seed, uint64_t ray_id, uint32_t emitter, uint32_t band, int ncalls) {
  if (g_faithful) { xoshiro_words(d->w, 2 * ncalls); return; }
  d->key[0] = (uint32_t)seed;
  d->key[1] = (uint32_t)(seed >> 32);
  d->ctr[0] = (uint32_t)ray_id;
  d->ctr[1] = (uint32_t)(ray_id >> 32);
  d->ctr[2] = emitter;
  for (int call = 0; call < ncalls; ++call) {
    d->ctr[3] = (band << 16) | (uint32_t)call;
    rthx_oracle_philox4x32_10(d->ctr, d->key, d->w + 4 * call);
  }
}

/* the two Philox calls of MULTI_BOUNCE event n: call# 2+2n and 3+2n */
static void draws_event(draws_t* d, uint64_t seed, uint64_t ray_id, uint32_t emitter, uint32_t band, int event) {
  if (g_faithful) { xoshiro_words(d->w, 4); return; }
  d->key[0] = (uint32_t)seed;
  d->key[1] = (uint32_t)(seed >> 32);
  d->ctr[0] = (uint32_t)ray_id;
  d->ctr[1] = (uint32_t)(ray_id >> 32);
  d->ctr[2] = emitter;
  for (int k = 0; k < 2; ++k) {
    d->ctr[3] = (band << 16) | (uint32_t)(2 + 2 * event + k);
    rthx_oracle_philox4x32_10(d->ctr, d->key, d->w + 4 * k);
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* geometry containers                                                                              */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
  int n;
  double vx[4], vy[4];
  double nx[4], ny[4]; /* the field the reference calls inwardNormals (they point outward) */
  double midx, midy;
  double volume;
  double bb[4]; /* min_x, max_x, min_y, max_y */
} poly_t;

typedef struct {
  double ox, oy, inv_cell, cell;
  int nx, ny;
  int* start; /* [nx*ny+1] CSR over buckets, bucket (i,j) -> i + j*nx */
  int* items;
} grid_t;

typedef struct {
  int n;
  poly_t* faces;
  grid_t grid;
} faceset_t;

typedef struct {
  const rthx_mesh* m;
  int n_coarse, n_cells, ns, N;
  faceset_t coarse;
  faceset_t* fine; /* [n_coarse] */
  int* em_cell;    /* [N] fine cell (global) of element e */
  int* em_wall;    /* [N] wall (0-based) for surface elements, -1 for volumes */
  int* em_coarse;  /* [N] */
} omesh_t;

/* calculateInwardNormal.jl:1-12 */
static void edge_normal(double x1, double y1, double x2, double y2, double midx, double midy, double* nx, double* ny) {
  const double ex = x2 - x1, ey = y2 - y1;
  double ax = ey, ay = -ex;
  const double len = sqrt(ax * ax + ay * ay);
  ax /= len; ay /= len;
  const double wmx = (x1 + x2) / 2, wmy = (y1 + y2) / 2;
  if (ax * (wmx - midx) + ay * (wmy - midy) < 0) { ax = -ax; ay = -ay; }
  *nx = ax; *ny = ay;
}

static void poly_finish(poly_t* p) {
  p->bb[0] = p->bb[2] = INFINITY;
  p->bb[1] = p->bb[3] = -INFINITY;
  for (int i = 0; i < p->n; ++i) {
    const int j = (i + 1) % p->n;
    edge_normal(p->vx[i], p->vy[i], p->vx[j], p->vy[j], p->midx, p->midy, &p->nx[i], &p->ny[i]);
    if (p->vx[i] < p->bb[0]) p->bb[0] = p->vx[i];
    if (p->vx[i] > p->bb[1]) p->bb[1] = p->vx[i];
    if (p->vy[i] < p->bb[2]) p->bb[2] = p->vy[i];
    if (p->vy[i] > p->bb[3]) p->bb[3] = p->vy[i];
  }
}

/* spatialAccelerations.jl:72-89 + :2-59 */
static int build_grid(faceset_t* fs) {
  grid_t* g = &fs->grid;
  double total = 0;
  for (int i = 0; i < fs->n; ++i) total += fs->faces[i].volume;
  const double cell = sqrt(total / fs->n) * 2.0;
  double min_x = INFINITY, min_y = INFINITY, max_x = -INFINITY, max_y = -INFINITY;
  for (int i = 0; i < fs->n; ++i) {
    const poly_t* p = &fs->faces[i];
    if (p->bb[0] < min_x) min_x = p->bb[0];
    if (p->bb[1] > max_x) max_x = p->bb[1];
    if (p->bb[2] < min_y) min_y = p->bb[2];
    if (p->bb[3] > max_y) max_y = p->bb[3];
  }
  const double pad = cell * 0.1;
  min_x -= pad; min_y -= pad; max_x += pad; max_y += pad;
  int nx = (int)ceil((max_x - min_x) / cell), ny = (int)ceil((max_y - min_y) / cell);
  if (nx < 1) nx = 1;
  if (ny < 1) ny = 1;
  g->ox = min_x; g->oy = min_y; g->cell = cell; g->inv_cell = 1.0 / cell; g->nx = nx; g->ny = ny;
  g->start = (int*)calloc((size_t)nx * ny + 1, sizeof(int));
  if (!g->start) return 1;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      int acc = 0;
      for (int b = 0; b < nx * ny; ++b) { const int c = g->start[b]; g->start[b] = acc; acc += c; }
      g->start[nx * ny] = acc;
      g->items = (int*)malloc(sizeof(int) * (size_t)(acc > 0 ? acc : 1));
      if (!g->items) return 1;
    }
    int* fill = NULL;
    if (pass == 1) { fill = (int*)malloc(sizeof(int) * (size_t)nx * ny); memcpy(fill, g->start, sizeof(int) * (size_t)nx * ny); }
    for (int f = 0; f < fs->n; ++f) { /* ascending face order => ascending order inside each bucket */
      const poly_t* p = &fs->faces[f];
      int si = (int)floor((p->bb[0] - min_x) / cell) + 1, ei = (int)ceil((p->bb[1] - min_x) / cell);
      int sj = (int)floor((p->bb[2] - min_y) / cell) + 1, ej = (int)ceil((p->bb[3] - min_y) / cell);
      if (si < 1) si = 1;
      if (sj < 1) sj = 1;
      if (ei > nx) ei = nx;
      if (ej > ny) ej = ny;
      for (int i = si; i <= ei; ++i)
        for (int j = sj; j <= ej; ++j) {
          const int b = (i - 1) + (j - 1) * nx;
          if (pass == 0) g->start[b]++; else g->items[fill[b]++] = f;
        }
    }
    free(fill);
  }
  return 0;
}

/* findFace2D.jl:77-101 */
static inline int point_in_polygon(double px, double py, const poly_t* f) {
  int inside = 0;
  int j = f->n - 1;
  for (int i = 0; i < f->n; ++i) {
    const double xi = f->vx[i], yi = f->vy[i], xj = f->vx[j], yj = f->vy[j];
    if ((yi > py) != (yj > py)) {
      const double slope = (xj - xi) / (yj - yi);
      const double ix = xi + slope * (py - yi);
      if (px < ix) inside = !inside;
    }
    j = i;
  }
  return inside;
}

/* findFace2D.jl:48-68 (grid :2-27, bbox prefilter :30-45); returns local face index or -1 (= nothing) */
static int find_face(const faceset_t* fs, double px, double py) {
  const grid_t* g = &fs->grid;
  const double rx = px - g->ox, ry = py - g->oy;
  const double fi = floor(rx * g->inv_cell), fj = floor(ry * g->inv_cell);
  if (fi >= 0 && fi < g->nx && fj >= 0 && fj < g->ny) {
    const int b = (int)fi + (int)fj * g->nx;
    for (int k = g->start[b]; k < g->start[b + 1]; ++k) {
      const int f = g->items[k];
      if (point_in_polygon(px, py, &fs->faces[f])) return f;
    }
  }
  for (int f = 0; f < fs->n; ++f) {
    const poly_t* p = &fs->faces[f];
    if (p->bb[0] <= px && px <= p->bb[1] && p->bb[2] <= py && py <= p->bb[3])
      if (point_in_polygon(px, py, p)) return f;
  }
  return -1;
}

/* distToSurface2D.jl:2-18: returns u, *idx = first index of the minimum (0-based); all-Inf -> (Inf, 0) */
static double dist_to_surface(double px, double py, double dx, double dy, const poly_t* f, int* idx) {
  double best = INFINITY;
  int bi = 0;
  for (int i = 0; i < f->n; ++i) {
    const double den = dx * f->nx[i] + dy * f->ny[i];
    double u;
    if (fabs(den) < 1e-10) u = INFINITY;
    else u = ((f->vx[i] - px) * f->nx[i] + (f->vy[i] - py) * f->ny[i]) / den;
    if (u <= 0) u = INFINITY;
    if (u < best) { best = u; bi = i; }
  }
  *idx = bi;
  return best;
}

static void omesh_free(omesh_t* o) {
  if (o->coarse.faces) free(o->coarse.faces);
  free(o->coarse.grid.start); free(o->coarse.grid.items);
  if (o->fine) {
    for (int c = 0; c < o->n_coarse; ++c) { free(o->fine[c].grid.start); free(o->fine[c].grid.items); free(o->fine[c].faces); }
    free(o->fine);
  }
  free(o->em_cell); free(o->em_wall); free(o->em_coarse);
}

static int omesh_build(omesh_t* o, const rthx_mesh* m) {
  memset(o, 0, sizeof(*o));
  o->m = m; o->n_coarse = m->n_coarse; o->n_cells = m->n_cells; o->ns = m->n_surfaces; o->N = m->n_surfaces + m->n_cells;
  o->coarse.n = m->n_coarse;
  o->coarse.faces = (poly_t*)calloc((size_t)m->n_coarse, sizeof(poly_t));
  o->fine = (faceset_t*)calloc((size_t)m->n_coarse, sizeof(faceset_t));
  o->em_cell = (int*)malloc(sizeof(int) * (size_t)o->N);
  o->em_wall = (int*)malloc(sizeof(int) * (size_t)o->N);
  o->em_coarse = (int*)malloc(sizeof(int) * (size_t)o->N);
  if (!o->coarse.faces || !o->fine || !o->em_cell || !o->em_wall || !o->em_coarse) return 1;
  for (int e = 0; e < o->N; ++e) o->em_cell[e] = -1;
  for (int c = 0; c < m->n_coarse; ++c) {
    poly_t* p = &o->coarse.faces[c];
    p->n = m->coarse_nv[c];
    if (p->n != 3 && p->n != 4) return 2;
    double sx = 0, sy = 0;
    for (int i = 0; i < p->n; ++i) { p->vx[i] = m->coarse_vx[c * 4 + i]; p->vy[i] = m->coarse_vy[c * 4 + i]; sx += p->vx[i]; sy += p->vy[i]; }
    p->midx = sx / p->n; p->midy = sy / p->n; /* PolyVolume2D.jl:9,103 */
    if (p->n == 4)
      p->volume = 0.5 * (p->vx[0] * (p->vy[1] - p->vy[2]) + p->vx[1] * (p->vy[2] - p->vy[0]) + p->vx[2] * (p->vy[0] - p->vy[1])) +
                  0.5 * (p->vx[2] * (p->vy[3] - p->vy[0]) + p->vx[3] * (p->vy[0] - p->vy[2]) + p->vx[0] * (p->vy[2] - p->vy[3]));
    else
      p->volume = 0.5 * (p->vx[0] * (p->vy[1] - p->vy[2]) + p->vx[1] * (p->vy[2] - p->vy[0]) + p->vx[2] * (p->vy[0] - p->vy[1]));
    poly_finish(p);
    faceset_t* fs = &o->fine[c];
    fs->n = m->fine_off[c + 1] - m->fine_off[c];
    if (fs->n <= 0) return 2;
    fs->faces = (poly_t*)calloc((size_t)fs->n, sizeof(poly_t));
    if (!fs->faces) return 1;
    for (int f = 0; f < fs->n; ++f) {
      const int g = m->fine_off[c] + f;
      poly_t* q = &fs->faces[f];
      q->n = m->cell_nv[g];
      if (q->n != 3 && q->n != 4) return 2;
      for (int i = 0; i < q->n; ++i) { q->vx[i] = m->cell_vx[g * 4 + i]; q->vy[i] = m->cell_vy[g * 4 + i]; }
      q->midx = m->cell_mid[g * 2]; q->midy = m->cell_mid[g * 2 + 1];
      q->volume = m->cell_volume[g];
      poly_finish(q);
      o->em_cell[o->ns + g] = g; o->em_wall[o->ns + g] = -1; o->em_coarse[o->ns + g] = c;
      for (int w = 0; w < q->n; ++w) {
        const int s = m->cell_surf_id[g * 4 + w];
        if (s >= 0) {
          if (s >= o->ns || o->em_cell[s] != -1) return 2;
          o->em_cell[s] = g; o->em_wall[s] = w; o->em_coarse[s] = c;
        }
      }
    }
    if (build_grid(fs)) return 1;
  }
  for (int e = 0; e < o->N; ++e) if (o->em_cell[e] < 0) return 2;
  if (build_grid(&o->coarse)) return 1;
  return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* emission                                                                                         */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { double px, py, dx, dy; } ray_t;

/* emitSurfaceRay2D.jl:1-27 + lambertSample2D.jl:1-11.  draws (rthx.h contract): w0 position (32-bit);
 * w1 cos-theta (Float32); w2 psi (Float32).  The free-path draw is (w4,w5), 52-bit. */
static ray_t emit_surface(const poly_t* face, int wall, double nudge, const draws_t* d) {
  const int j = (wall + 1) % face->n;
  const double p1x = face->vx[wall], p1y = face->vy[wall], p2x = face->vx[j], p2y = face->vy[j];
  const double R = rthx_oracle_u32(d->w[0]);
  double px = p1x + (p2x - p1x) * R, py = p1y + (p2y - p1y) * R;
  px = px + (face->midx - px) * nudge;
  py = py + (face->midy - py) * nudge;
  /* lambertSample2D: Float32 variates, Float32 sqrt and square, the rest in Float64 */
  const float R_angle1 = rthx_oracle_u23(d->w[1]);
  const float cosTheta = sqrtf(R_angle1);
  const float cos2 = cosTheta * cosTheta;
  const double sinTheta = sqrt(1.0 - (double)cos2);
  const double psi = (2.0 * M_PI) * (double)rthx_oracle_u23(d->w[2]);
  const double xdir = sinTheta * cos(psi);
  const double zdir = (double)cosTheta;
  const double ex = p2x - p1x, ey = p2y - p1y;
  const double len = sqrt(ex * ex + ey * ey);
  const double xlx = ex / len, xly = ey / len; /* xVecLocal */
  const double ylx = -xly, yly = xlx;          /* yVecLocal */
  ray_t r;
  r.px = px; r.py = py;
  r.dx = xlx * xdir + ylx * zdir;
  r.dy = xly * xdir + yly * zdir;
  return r;
}

/* emitVolumeRay2D.jl:1-33.  draws (rthx.h contract): w0 R_1, w1 R_2, w2 triangle selector (quads only; the slot is
 * skipped for triangles), w3 phi — 32-bit each; (w4,w5) theta, 52-bit.  The free-path draw is (w6,w7), 52-bit. */
static ray_t emit_volume(const poly_t* face, double nudge, const draws_t* d) {
  const double R_1 = rthx_oracle_u32(d->w[0]), R_2 = rthx_oracle_u32(d->w[1]);
  const double sqrt_R1 = sqrt(R_1);
  double px, py;
  const double Ax = face->vx[0], Ay = face->vy[0], Bx = face->vx[1], By = face->vy[1], Cx = face->vx[2], Cy = face->vy[2];
  if (face->n == 4) {
    const double Dx = face->vx[3], Dy = face->vy[3];
    const double sel = rthx_oracle_u32(d->w[2]);
    if (sel < 0.5 * (Ax * (By - Cy) + Bx * (Cy - Ay) + Cx * (Ay - By)) / face->volume) {
      px = (1 - sqrt_R1) * Ax + sqrt_R1 * (1 - R_2) * Bx + sqrt_R1 * R_2 * Cx;
      py = (1 - sqrt_R1) * Ay + sqrt_R1 * (1 - R_2) * By + sqrt_R1 * R_2 * Cy;
    } else {
      px = (1 - sqrt_R1) * Cx + sqrt_R1 * (1 - R_2) * Dx + sqrt_R1 * R_2 * Ax;
      py = (1 - sqrt_R1) * Cy + sqrt_R1 * (1 - R_2) * Dy + sqrt_R1 * R_2 * Ay;
    }
  } else {
    px = (1 - sqrt_R1) * Ax + sqrt_R1 * (1 - R_2) * Bx + sqrt_R1 * R_2 * Cx;
    py = (1 - sqrt_R1) * Ay + sqrt_R1 * (1 - R_2) * By + sqrt_R1 * R_2 * Cy;
  }
  px = px + (face->midx - px) * nudge;
  py = py + (face->midy - py) * nudge;
  const double theta = acos(1 - 2 * rthx_oracle_u52(d->w[4], d->w[5]));
  const double phi = (2.0 * M_PI) * rthx_oracle_u32(d->w[3]);
  ray_t r;
  r.px = px; r.py = py;
  r.dx = sin(theta) * cos(phi);
  r.dy = cos(theta);
  return r;
}

/* ------------------------------------------------------------------------------------------------ */
/* traversal                                                                                        */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int ok, coarse, fine, wall; double px, py; int crossings; } hit_t; /* wall: 0 = gas, 1.. = wall */

/* traceRayUniform, traceRay.jl:20-70.  R_S is the uniform behind `-log(rand())/beta` (:25). */
static hit_t trace_uniform(const omesh_t* o, ray_t r, double beta, double nudge, int coarse, double R_S) {
  hit_t h; memset(&h, 0, sizeof(h));
  double px = r.px, py = r.py;
  double S = beta > 0 ? -log(R_S) / beta : INFINITY;
  for (int it = 0; it < 10000; ++it) {
    const poly_t* cf = &o->coarse.faces[coarse];
    int k;
    const double u = dist_to_surface(px, py, r.dx, r.dy, cf, &k);
    if (S < u) {
      px = px + (S - nudge) * r.dx; py = py + (S - nudge) * r.dy;
      const int f = find_face(&o->fine[coarse], px, py);
      if (f < 0) return h;
      h.ok = 1; h.coarse = coarse; h.fine = f; h.wall = 0; h.px = px; h.py = py;
      return h;
    } else if (o->m->coarse_solid[coarse * 4 + k]) {
      px = px + (u - nudge) * r.dx; py = py + (u - nudge) * r.dy;
      const int f = find_face(&o->fine[coarse], px, py);
      if (f < 0) return h;
      int kw;
      (void)dist_to_surface(px, py, r.dx, r.dy, &o->fine[coarse].faces[f], &kw);
      h.ok = 1; h.coarse = coarse; h.fine = f; h.wall = kw + 1; h.px = px; h.py = py;
      return h;
    } else {
      px = px + (u + nudge) * r.dx; py = py + (u + nudge) * r.dy;
      S -= u;
      const int nc = find_face(&o->coarse, px, py);
      if (nc < 0) return h;
      coarse = nc;
      h.crossings++;
    }
  }
  return h;
}

/* traceRayVariable, traceRay.jl:73-147.  R_S is the uniform behind `target_tau = -log(rand())` (:79). */
static hit_t trace_variable(const omesh_t* o, ray_t r, const double* beta_band, double nudge, int coarse, double R_S) {
  hit_t h; memset(&h, 0, sizeof(h));
  double px = r.px, py = r.py;
  const double target_tau = -log(R_S);
  double acc = 0.0;
  for (int it = 0; it < 10000; ++it) {
    const poly_t* cf = &o->coarse.faces[coarse];
    int k;
    const double u = dist_to_surface(px, py, r.dx, r.dy, cf, &k);
    int f = find_face(&o->fine[coarse], px, py);
    if (f < 0) return h;
    const double local_beta = beta_band[o->m->fine_off[coarse] + f];
    const double tau_b = local_beta * u;
    if (acc + tau_b >= target_tau) {
      const double S = (target_tau - acc) / local_beta;
      px = px + (S - nudge) * r.dx; py = py + (S - nudge) * r.dy;
      f = find_face(&o->fine[coarse], px, py);
      if (f < 0) return h;
      h.ok = 1; h.coarse = coarse; h.fine = f; h.wall = 0; h.px = px; h.py = py;
      return h;
    } else if (o->m->coarse_solid[coarse * 4 + k]) {
      px = px + (u - nudge) * r.dx; py = py + (u - nudge) * r.dy;
      f = find_face(&o->fine[coarse], px, py);
      if (f < 0) return h;
      int kw;
      (void)dist_to_surface(px, py, r.dx, r.dy, &o->fine[coarse].faces[f], &kw);
      h.ok = 1; h.coarse = coarse; h.fine = f; h.wall = kw + 1; h.px = px; h.py = py;
      return h;
    } else {
      px = px + (u + nudge) * r.dx; py = py + (u + nudge) * r.dy;
      acc += tau_b;
      const int nc = find_face(&o->coarse, px, py);
      if (nc < 0) return h;
      coarse = nc;
      h.crossings++;
    }
  }
  return h;
}

/* one ray of element e: emission, trace, global index (getGlobalIndex2D.jl:1-14). Returns absorber or -1. */
static int shoot(const omesh_t* o, const double* beta_all, const rthx_trace_args* a, int e, int band, uint64_t ray_id,
                 ray_t* ray_out, hit_t* hit_out) {
  const rthx_mesh* m = o->m;
  const int g = o->em_cell[e], c = o->em_coarse[e];
  const poly_t* cell = &o->fine[c].faces[g - m->fine_off[c]];
  draws_t d;
  ray_t r;
  double R_S;
  if (e < o->ns) {
    draws_init(&d, a->seed, ray_id, (uint32_t)e, (uint32_t)band, 2);
    r = emit_surface(cell, o->em_wall[e], a->nudge, &d);
    R_S = rthx_oracle_u52(d.w[4], d.w[5]);
  } else {
    draws_init(&d, a->seed, ray_id, (uint32_t)e, (uint32_t)band, 2);
    r = emit_volume(cell, a->nudge, &d);
    R_S = rthx_oracle_u52(d.w[6], d.w[7]);
  }
  const double* beta_band = beta_all + (size_t)band * m->n_cells;
  const int uniform = m->uniform_beta[band] > -0.1; /* traceRay.jl:4-12: uniform bins use beta of fine_mesh[1][1] */
  hit_t h = uniform ? trace_uniform(o, r, beta_band[0], a->nudge, c, R_S) : trace_variable(o, r, beta_band, a->nudge, c, R_S);
  if (ray_out) *ray_out = r;
  int crossings = h.crossings;
  /* RTHX_MULTI_BOUNCE: traceSingleRay.jl:7-81 without the re-emission branches — absorb, scatter or reflect at every
   * interaction until the ray is absorbed. */
  for (int event = 0;; ++event) {
    if (!h.ok) { if (hit_out) { *hit_out = h; hit_out->crossings = crossings; } return -1; }
    const int gh = m->fine_off[h.coarse] + h.fine;
    const int absorber = h.wall > 0 ? m->cell_surf_id[gh * 4 + (h.wall - 1)] : o->ns + gh; /* getGlobalIndex2D.jl:1-14 */
    if (hit_out) { *hit_out = h; hit_out->crossings = crossings; }
    if (a->mode == RTHX_FIRST_INTERACTION || absorber < 0) return absorber; /* -1: that fine wall is not solid */
    if (event >= 16000) return -1;
    draws_event(&d, a->seed, ray_id, (uint32_t)e, (uint32_t)band, event);
    if (event >= 1000 && rthx_oracle_u32(d.w[3]) > 0.8) return -1; /* Russian roulette, traceSingleRay.jl:11 */
    const double dec = rthx_oracle_u32(d.w[0]);
    ray_t r2;
    r2.px = h.px; r2.py = h.py; /* origin = end_point (:45,:64) */
    if (h.wall == 0) {
      const size_t ib = (size_t)band * m->n_cells + gh;
      const double beta = m->kappa[ib] + m->sigma_s[ib];
      const double omega = beta > 0 ? m->sigma_s[ib] / beta : 0.0;
      if (!(dec < omega)) return absorber; /* :61 */
      /* isotropicScatter2D.jl:1-4 */
      const double theta = acos(2 * rthx_oracle_u52(d.w[4], d.w[5]) - 1);
      const double phi = (2.0 * M_PI) * rthx_oracle_u32(d.w[1]);
      r2.dx = sin(theta) * cos(phi);
      r2.dy = cos(theta);
    } else {
      const double eps = m->epsilon ? m->epsilon[(size_t)band * o->ns + absorber] : 1.0;
      if (dec < eps) return absorber; /* :28 */
      const poly_t* cell = &o->fine[h.coarse].faces[h.fine];
      const double nx = -cell->nx[h.wall - 1], ny = -cell->ny[h.wall - 1]; /* inward normal of the wall */
      if (a->mode == RTHX_MULTI_BOUNCE_SPECULAR) {
        const double dn = r.dx * nx + r.dy * ny;
        r2.dx = r.dx - 2 * dn * nx;
        r2.dy = r.dy - 2 * dn * ny;
      } else {
        /* sampleReflectionDirection2D.jl:5-16 with lambertSample2D.jl:1-11: x-axis (n.y, -n.x), y-axis n */
        const float cosTheta = sqrtf(rthx_oracle_u23(d.w[2]));
        const float cos2 = cosTheta * cosTheta;
        const double sinTheta = sqrt(1.0 - (double)cos2);
        const double xdir = sinTheta * cos((2.0 * M_PI) * (double)rthx_oracle_u23(d.w[1]));
        const double zdir = (double)cosTheta;
        r2.dx = ny * xdir + nx * zdir;
        r2.dy = -nx * xdir + ny * zdir;
      }
    }
    r = r2;
    R_S = rthx_oracle_u52(d.w[6], d.w[7]);
    h = uniform ? trace_uniform(o, r, beta_band[0], a->nudge, h.coarse, R_S) : trace_variable(o, r, beta_band, a->nudge, h.coarse, R_S);
    crossings += h.crossings;
  }
}

static int check_args(const rthx_mesh* m, const rthx_trace_args* a) {
  if (!m || !a || a->rays_per_emitter < 0 || a->n_bins < 1 || !a->bins) return 1;
  for (int b = 0; b < a->n_bins; ++b) if (a->bins[b] < 0 || a->bins[b] >= m->n_bands) return 1;
  if (a->emitter_world < 1 || a->emitter_rank < 0 || a->emitter_rank >= a->emitter_world) return 1;
  if (a->mode != RTHX_FIRST_INTERACTION && a->mode != RTHX_MULTI_BOUNCE && a->mode != RTHX_MULTI_BOUNCE_SPECULAR) return 1;
  return 0;
}

int rthx_oracle_trace(const rthx_mesh* m, const rthx_trace_args* a, uint64_t* counts, uint64_t* lost,
                      rthx_rec_out* rec, int n_threads, rthx_oracle_stats* st) {
  if (check_args(m, a) || !counts) return RTHX_ERR_ARG;
  omesh_t o;
  const int rc = omesh_build(&o, m);
  if (rc) { omesh_free(&o); return rc == 1 ? RTHX_ERR_NOMEM : RTHX_ERR_ARG; }
  const int N = o.N;
  double* beta_all = (double*)malloc(sizeof(double) * (size_t)m->n_bands * m->n_cells);
  for (size_t i = 0; i < (size_t)m->n_bands * m->n_cells; ++i) beta_all[i] = m->kappa[i] + m->sigma_s[i];
  memset(counts, 0, sizeof(uint64_t) * (size_t)a->n_bins * N * N);
  if (lost) memset(lost, 0, sizeof(uint64_t) * (size_t)a->n_bins * N);
  /* recorder slots: deterministic (sorted emitter, ray id) order, compacted below */
  int* rec_slot = (int*)malloc(sizeof(int) * (size_t)N);
  for (int e = 0; e < N; ++e) rec_slot[e] = -1;
  int n_rec = 0;
  double* rbuf = NULL; unsigned char* rvalid = NULL;
  if (rec && a->n_rec_ids > 0) {
    for (int i = 0; i < a->n_rec_ids; ++i) if (a->rec_ids[i] >= 0 && a->rec_ids[i] < N) rec_slot[a->rec_ids[i]] = 0;
    for (int e = 0; e < N; ++e) if (rec_slot[e] == 0) rec_slot[e] = n_rec++;
    rbuf = (double*)malloc(sizeof(double) * 4 * (size_t)n_rec * (size_t)a->rays_per_emitter + 8);
    rvalid = (unsigned char*)calloc((size_t)n_rec * (size_t)a->rays_per_emitter + 1, 1);
  }
  uint64_t c_sg = 0, c_sw = 0, c_vg = 0, c_vw = 0, c_cross = 0, c_lost = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#else
  (void)n_threads;
#endif
  const int faithful = g_faithful;
#ifdef _OPENMP
  const double t_loop0 = omp_get_wtime();
#endif
  /* contiguous emitter ranges per thread, like parallelRayTracing.jl:83-91,102 */
#pragma omp parallel reduction(+ : c_sg, c_sw, c_vg, c_vw, c_cross, c_lost)
  {
  rowmap_t rm = {0, 0, 0, 0};
  if (faithful) {
    rowmap_init(&rm);
#ifdef _OPENMP
    xoshiro_seed(a->seed, (uint64_t)omp_get_thread_num());
#else
    xoshiro_seed(a->seed, 0);
#endif
  }
#pragma omp for schedule(static)
  for (int e = 0; e < N; ++e) {
    if (e % a->emitter_world != a->emitter_rank) continue;
    for (int bi = 0; bi < a->n_bins; ++bi) {
      const int band = a->bins[bi];
      uint64_t* row = counts + ((size_t)bi * N + e) * N;
      const int recording = rbuf && rec_slot[e] >= 0 && band == a->rec_bin;
      uint64_t nlost = 0;
      for (int64_t i = 0; i < a->rays_per_emitter; ++i) {
        ray_t r; hit_t h;
        const int ab = shoot(&o, beta_all, a, e, band, (uint64_t)(a->ray_id_offset + i), &r, &h);
        c_cross += (uint64_t)h.crossings;
        if (ab < 0) { nlost++; continue; }
        if (faithful) rowmap_add(&rm, ab); else row[ab]++;
        if (e < o.ns) { if (h.wall) c_sw++; else c_sg++; } else { if (h.wall) c_vw++; else c_vg++; }
        if (recording) {
          const size_t s = (size_t)rec_slot[e] * (size_t)a->rays_per_emitter + (size_t)i;
          rbuf[4 * s] = r.px; rbuf[4 * s + 1] = r.py; rbuf[4 * s + 2] = h.px; rbuf[4 * s + 3] = h.py;
          rvalid[s] = 1;
        }
      }
      if (faithful) {                      /* flush the emitter's tallies (parallelRayTracing.jl:144-146) */
        for (uint32_t k = 0; k < rm.cap; ++k) if (rm.keys[k] != -1) row[rm.keys[k]] += rm.vals[k];
        rowmap_clear(&rm);
      }
      if (lost) lost[(size_t)bi * N + e] = nlost;
      c_lost += nlost;
    }
  }
  if (faithful) rowmap_free(&rm);
  }
#ifdef _OPENMP
  const double t_loop1 = omp_get_wtime();
#endif
  if (rec) {
    rec->n_recorded = 0;
    if (rbuf) {
      const size_t tot = (size_t)n_rec * (size_t)a->rays_per_emitter;
      for (size_t s = 0; s < tot; ++s) {
        if (!rvalid[s]) continue;
        if (rec->n_recorded >= rec->capacity) break;
        const int64_t k = rec->n_recorded++;
        rec->origins[2 * k] = rbuf[4 * s]; rec->origins[2 * k + 1] = rbuf[4 * s + 1];
        rec->endpoints[2 * k] = rbuf[4 * s + 2]; rec->endpoints[2 * k + 1] = rbuf[4 * s + 3];
      }
    }
  }
  if (st) {
    st->n_surface_gas = c_sg; st->n_surface_wall = c_sw; st->n_volume_gas = c_vg; st->n_volume_wall = c_vw;
    st->n_crossings = c_cross; st->n_lost = c_lost;
#ifdef _OPENMP
    st->n_threads = omp_get_max_threads();
    st->loop_seconds = t_loop1 - t_loop0;
#else
    st->n_threads = 1;
    st->loop_seconds = 0.0;
#endif
  }
  free(rbuf); free(rvalid); free(rec_slot); free(beta_all);
  omesh_free(&o);
  return RTHX_OK;
}

/* Single-ray probe for unit tests: out = {px, py, dx, dy, hit_x, hit_y}, returns absorber element or -1. */
int rthx_oracle_shoot(const rthx_mesh* m, const rthx_trace_args* a, int emitter, int band, uint64_t ray_id, double out[6]) {
  if (check_args(m, a)) return -2;
  omesh_t o;
  if (omesh_build(&o, m)) { omesh_free(&o); return -2; }
  if (emitter < 0 || emitter >= o.N) { omesh_free(&o); return -2; }
  double* beta_all = (double*)malloc(sizeof(double) * (size_t)m->n_bands * m->n_cells);
  for (size_t i = 0; i < (size_t)m->n_bands * m->n_cells; ++i) beta_all[i] = m->kappa[i] + m->sigma_s[i];
  ray_t r; hit_t h;
  const int ab = shoot(&o, beta_all, a, emitter, band, ray_id, &r, &h);
  out[0] = r.px; out[1] = r.py; out[2] = r.dx; out[3] = r.dy; out[4] = h.px; out[5] = h.py;
  free(beta_all);
  omesh_free(&o);
  return ab;
}

/* Bulk emission probe (sampler statistics tests): out[i] = {px,py,dx,dy} for ray ids 0..n-1 of `emitter`. */
int rthx_oracle_emit(const rthx_mesh* m, const rthx_trace_args* a, int emitter, int band, int64_t n, double* out) {
  if (check_args(m, a)) return RTHX_ERR_ARG;
  omesh_t o;
  if (omesh_build(&o, m)) { omesh_free(&o); return RTHX_ERR_ARG; }
  if (emitter < 0 || emitter >= o.N) { omesh_free(&o); return RTHX_ERR_ARG; }
  const int g = o.em_cell[emitter], c = o.em_coarse[emitter];
  const poly_t* cell = &o.fine[c].faces[g - m->fine_off[c]];
  for (int64_t i = 0; i < n; ++i) {
    draws_t d;
    ray_t r;
    draws_init(&d, a->seed, (uint64_t)(a->ray_id_offset + i), (uint32_t)emitter, (uint32_t)band, 2);
    r = emitter < o.ns ? emit_surface(cell, o.em_wall[emitter], a->nudge, &d) : emit_volume(cell, a->nudge, &d);
    out[4 * i] = r.px; out[4 * i + 1] = r.py; out[4 * i + 2] = r.dx; out[4 * i + 3] = r.dy;
  }
  omesh_free(&o);
  return RTHX_OK;
}

/* Point-location probe: face set -1 = coarse mesh, c >= 0 = fine mesh of coarse face c. */
int rthx_oracle_find_face(const rthx_mesh* m, int set, double px, double py) {
  omesh_t o;
  if (omesh_build(&o, m)) { omesh_free(&o); return -2; }
  int f = -2;
  if (set == -1) f = find_face(&o.coarse, px, py);
  else if (set >= 0 && set < o.n_coarse) f = find_face(&o.fine[set], px, py);
  omesh_free(&o);
  return f;
}

/* distToSurface2D probe on an ad-hoc polygon: returns u, *idx 0-based. */
double rthx_oracle_dist_to_surface(int n, const double* vx, const double* vy, double px, double py, double dx, double dy, int* idx) {
  poly_t p; memset(&p, 0, sizeof(p));
  p.n = n;
  double sx = 0, sy = 0;
  for (int i = 0; i < n; ++i) { p.vx[i] = vx[i]; p.vy[i] = vy[i]; sx += vx[i]; sy += vy[i]; }
  p.midx = sx / n; p.midy = sy / n;
  poly_finish(&p);
  return dist_to_surface(px, py, dx, dy, &p, idx);
}
