// rthx_grid.h — host-only tables of the generic locator: polygons with their edge normals, the bucket grid of a face set and the
// per-polygon records the kernels read.  Plain C++ (no device code): included by rthx_api.cu, and by tests/grid_check.cpp, which
// replays the device's locator on the CPU against a brute-force scan.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "rthx_internal.h"

namespace rthx {

struct Poly { int n; double vx[4], vy[4], nx[4], ny[4], midx, midy, volume, bb[4]; };

inline void edge_normal(double x1, double y1, double x2, double y2, double midx, double midy, double* nx, double* ny) {
  const double ex = x2 - x1, ey = y2 - y1;
  double ax = ey, ay = -ex;
  const double len = std::sqrt(ax * ax + ay * ay);
  ax /= len; ay /= len;
  const double wmx = (x1 + x2) / 2, wmy = (y1 + y2) / 2;
  if (ax * (wmx - midx) + ay * (wmy - midy) < 0) { ax = -ax; ay = -ay; }
  *nx = ax; *ny = ay;
}

inline void poly_finish(Poly& p) {
  p.bb[0] = p.bb[2] = INFINITY; p.bb[1] = p.bb[3] = -INFINITY;
  for (int i = 0; i < p.n; ++i) {
    const int j = (i + 1) % p.n;
    edge_normal(p.vx[i], p.vy[i], p.vx[j], p.vy[j], p.midx, p.midy, &p.nx[i], &p.ny[i]);
    p.bb[0] = std::min(p.bb[0], p.vx[i]); p.bb[1] = std::max(p.bb[1], p.vx[i]);
    p.bb[2] = std::min(p.bb[2], p.vy[i]); p.bb[3] = std::max(p.bb[3], p.vy[i]);
  }
  for (int i = p.n; i < 4; ++i) { p.vx[i] = p.vy[i] = p.nx[i] = p.ny[i] = 0.0; }
}

// Bucket grid of one face set for the generic locator (the role of spatialAccelerations.jl:72-89 + :2-59).  Which face contains a
// point does not depend on the grid, so the device's grid is built for the hardware rather than copied from the reference (square
// buckets of 2 sqrt(mean area), ~9 bounding-box candidates each): the generic kernel is bound by the L1 data pipe — one 32-byte
// sector per lane and load instruction, ~19 per ray with round 2's first tables (ncu: l1tex__data_pipe_lsu_wavefronts 98.7 %) — so
// the tables are shaped to answer most locations from ONE sector:
//   * buckets of 1/RTHX_GRID_FINE (default 1/12 for sets of up to 4096 faces, else 1/8) of the mean bounding-box extent per axis
//     (anisotropic: a 1000:1 slab keeps its candidate count), at most RTHX_GRID_PER_FACE n + 256 of them (default 200).  Measured on cfg3 /
//     cfg5 (tools/gpu_generic_sweep.sh): 1/4 2.49e10 / 4.1e9, 1/8 2.87e10 / 4.7e9, 1/12 3.00e10 / 5.8e9, 1/16 3.08e10 / 5.7e9 rays/s;
//   * candidates by the separating-axis test polygon <-> bucket rectangle (not bounding-box overlap: triangles and skewed cells
//     stop leaking into their neighbours' buckets);
//   * a bucket that lies wholly inside its only candidate is marked SOLE: the locator returns the face without any vertex test
//     ((1 - 1/12)^2 = 84 % of the buckets of a regular mesh, 77 % at 1/8);
//   * one 16-byte entry per bucket {code, a, b, c}: code -1 empty, 0 sole (a = face), k = 1..3 candidates a, b, c (crossing-number
//     test in ascending face order: "first PIP hit" as in the reference), k > 3 candidates cand[a .. a + k).
// Non-convex polygons (the reference never builds one) fall back to bounding-box candidates and are never sole.
// Margins: a polygon is dropped from a bucket only if an edge separates them by more than 1e-9 of the bucket size, the bucket is
// sole only if it is inside by twice that; the rectangle itself is widened by 1e-9 buckets (rounding of the device's bucket index).
struct GridCfg { int fine = 0; int per_face = 200; };
inline GridCfg grid_cfg(int n_faces) {       // A/B knobs (tools/gpu_generic_sweep.sh)
  GridCfg g;
  // 1/12 buys 4 % (cfg3) to 20 % (cfg5) of throughput over 1/8 and costs 2.2x the buckets: taken for sets whose tables stay small
  // (<= 4096 faces: 13 MB); larger sets use 1/8 — on cfg3 (10 201 cells) the tables are then 10 MB and 25 ms of host work instead of
  // 24 MB and 50 ms, which a single 1e9-ray trace (33 ms) would not earn back
  g.fine = n_faces <= 4096 ? 12 : 8;
  if (const char* ev = std::getenv("RTHX_GRID_FINE")) g.fine = std::min(32, std::max(1, std::atoi(ev)));
  if (const char* ev = std::getenv("RTHX_GRID_PER_FACE")) g.per_face = std::min(1024, std::max(1, std::atoi(ev)));
  return g;
}

inline void build_grid(const Poly* faces, int n, int poly_base, FaceSetDev& fs, std::vector<int32_t>& ent, std::vector<int32_t>& cand) {
  double min_x = INFINITY, min_y = INFINITY, max_x = -INFINITY, max_y = -INFINITY, ex = 0, ey = 0, total = 0;
  for (int i = 0; i < n; ++i) {
    min_x = std::min(min_x, faces[i].bb[0]); max_x = std::max(max_x, faces[i].bb[1]);
    min_y = std::min(min_y, faces[i].bb[2]); max_y = std::max(max_y, faces[i].bb[3]);
    ex += faces[i].bb[1] - faces[i].bb[0]; ey += faces[i].bb[3] - faces[i].bb[2];
    total += faces[i].volume;
  }
  if (n == 0) { min_x = min_y = 0; max_x = max_y = 1; }
  const double iso = n ? std::sqrt(std::fabs(total) / n) : 1.0;
  const GridCfg cfg = grid_cfg(n);
  const double fine = (double)cfg.fine;
  double sx = n ? ex / n / fine : 1.0, sy = n ? ey / n / fine : 1.0;
  if (!(sx > 0) || !std::isfinite(sx)) sx = iso > 0 ? iso : 1.0;
  if (!(sy > 0) || !std::isfinite(sy)) sy = iso > 0 ? iso : 1.0;
  const double max_buckets = (double)cfg.per_face * n + 256.0;
  int nx = 1, ny = 1;
  double ox = 0, oy = 0;
  for (int attempt = 0;; ++attempt) {
    ox = min_x - 0.1 * sx; oy = min_y - 0.1 * sy;
    const double fx = std::ceil(((max_x + 0.1 * sx) - ox) / sx), fy = std::ceil(((max_y + 0.1 * sy) - oy) / sy);
    if (!(fx * fy <= max_buckets) && attempt < 200) { sx *= 1.25; sy *= 1.25; continue; }
    nx = std::max(1, (int)fx); ny = std::max(1, (int)fy);
    break;
  }
  const double inv_cx = 1.0 / sx, inv_cy = 1.0 / sy;
  const double m_sep = 1e-9 * std::max(sx, sy), m_in = 2.0 * m_sep;
  // convexity (all turns of one sign)
  std::vector<uint8_t> convex((size_t)n, 0);
  for (int f = 0; f < n; ++f) {
    const Poly& q = faces[f];
    int pos = 0, neg = 0;
    for (int i = 0; i < q.n; ++i) {
      const int j = (i + 1) % q.n, k = (i + 2) % q.n;
      const double cr = (q.vx[j] - q.vx[i]) * (q.vy[k] - q.vy[j]) - (q.vy[j] - q.vy[i]) * (q.vx[k] - q.vx[j]);
      if (cr > 0) ++pos; else if (cr < 0) ++neg;
    }
    convex[f] = (q.n >= 3 && (pos == 0 || neg == 0) && pos + neg > 0) ? 1 : 0;
  }
  auto range = [&](const Poly& f, int& i0, int& i1, int& j0, int& j1) {
    const double slack = 1e-9;
    i0 = std::max(0, (int)std::floor((f.bb[0] - ox) * inv_cx - slack)); i1 = std::min(nx - 1, (int)std::floor((f.bb[1] - ox) * inv_cx + slack));
    j0 = std::max(0, (int)std::floor((f.bb[2] - oy) * inv_cy - slack)); j1 = std::min(ny - 1, (int)std::floor((f.bb[3] - oy) * inv_cy + slack));
  };
  const bool timing = std::getenv("RTHX_GRID_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto t = std::chrono::steady_clock::now();
    std::fprintf(stderr, "build_grid(%d faces): %-12s %8.3f ms\n", n, what, std::chrono::duration<double, std::milli>(t - t_prev).count());
    t_prev = t;
  };
  // Three passes, each split over a few host threads (cfg3: 10 201 cells, 1.5e6 buckets, 3.4e6 face-bucket pairs — 80 ms single-threaded
  // with two relation passes, ~20 ms now):
  //   1. (faces)   relation of face f to every bucket (i, j) of its bounding-box range — 0 apart, 1 candidate, 2 the bucket lies inside
  //                the face — from the signed distances of the rectangle's extreme corners to every edge line (outward positive; the
  //                extremes follow from the signs of the normal), the y part hoisted out of the row; bucket sizes by atomic increments;
  //   2. (faces)   fill through atomic cursors;
  //   3. (buckets) every bucket's candidates sorted into ascending face order (the fill order depends on the thread timing, the
  //                tables do not) and the 16-byte entry written.
  const size_t nb = (size_t)nx * ny;
  std::vector<size_t> off((size_t)n + 1, 0);
  for (int f = 0; f < n; ++f) {
    int i0, i1, j0, j1; range(faces[f], i0, i1, j0, j1);
    off[(size_t)f + 1] = off[f] + (size_t)std::max(0, i1 - i0 + 1) * (size_t)std::max(0, j1 - j0 + 1);
  }
  int n_thr = 1;
  if (off[n] > (size_t)1 << 18) n_thr = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
  if (const char* ev = std::getenv("RTHX_GRID_THREADS")) n_thr = std::min(64, std::max(1, std::atoi(ev)));
  auto parallel = [&](size_t count, auto&& body) {      // body(begin, end) over [0, count) in n_thr contiguous pieces
    if (n_thr <= 1 || count < 2) { body((size_t)0, count); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_thr; ++t) pool.emplace_back([&, t] { body(count * t / n_thr, count * (t + 1) / n_thr); });
    for (auto& th : pool) th.join();
  };
  std::vector<uint8_t> code(off[n]);
  std::vector<int32_t> cnt(nb + 1, 0);
  std::vector<uint8_t> sole(nb, 0);
  parallel((size_t)n, [&](size_t f_begin, size_t f_end) {
    for (size_t f = f_begin; f < f_end; ++f) {
      const Poly& q = faces[f];
      int i0, i1, j0, j1; range(q, i0, i1, j0, j1);
      if (i1 < i0 || j1 < j0) continue;
      uint8_t* out = code.data() + off[f];
      double ea[4], eh[4], eb[4];
      for (int e = 0; e < q.n; ++e) { ea[e] = q.nx[e]; eb[e] = q.ny[e]; eh[e] = q.vx[e] * q.nx[e] + q.vy[e] * q.ny[e]; }
      for (int j = j0; j <= j1; ++j) {
        const double y0 = oy + (j - 1e-9) * sy, y1 = oy + (j + 1 + 1e-9) * sy;
        double ylo[4], yhi[4];
        for (int e = 0; e < q.n; ++e) { ylo[e] = (eb[e] >= 0 ? y0 : y1) * eb[e] - eh[e]; yhi[e] = (eb[e] >= 0 ? y1 : y0) * eb[e] - eh[e]; }
        for (int i = i0; i <= i1; ++i) {
          int r = 1;                                     // non-convex polygons: bounding-box candidates, never sole
          if (convex[f]) {
            const double x0 = ox + (i - 1e-9) * sx, x1 = ox + (i + 1 + 1e-9) * sx;
            r = 2;
            for (int e = 0; e < q.n; ++e) {
              const double lo = (ea[e] >= 0 ? x0 : x1) * ea[e] + ylo[e], hi = (ea[e] >= 0 ? x1 : x0) * ea[e] + yhi[e];
              if (lo > m_sep) { r = 0; break; }
              if (!(hi < -m_in)) r = 1;
            }
          }
          *out++ = (uint8_t)r;
          if (r) {
            const size_t b = (size_t)i + (size_t)j * nx;
            __atomic_fetch_add(&cnt[b + 1], 1, __ATOMIC_RELAXED);
            if (r == 2) __atomic_store_n(&sole[b], (uint8_t)1, __ATOMIC_RELAXED);
          }
        }
      }
    }
  });
  lap("classify");
  for (size_t b = 0; b < nb; ++b) cnt[b + 1] += cnt[b];
  std::vector<int32_t> items((size_t)cnt[nb]);
  // cnt[b] is the cursor of bucket b: its start before the fill, its end afterwards (start of b = end of b - 1)
  parallel((size_t)n, [&](size_t f_begin, size_t f_end) {
    for (size_t f = f_begin; f < f_end; ++f) {
      int i0, i1, j0, j1; range(faces[f], i0, i1, j0, j1);
      const uint8_t* c = code.data() + off[f];
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i)
          if (*c++) items[(size_t)__atomic_fetch_add(&cnt[(size_t)i + (size_t)j * nx], 1, __ATOMIC_RELAXED)] = (int32_t)f;
    }
  });
  lap("fill");
  // overflow lists (more than three candidates) are laid out in bucket order: offsets by a serial scan
  std::vector<int32_t> cand_off;
  size_t n_over = 0;
  {
    size_t total = 0;
    auto size_of = [&](size_t b) { return cnt[b] - (b ? cnt[b - 1] : 0); };
    for (size_t b = 0; b < nb; ++b) if (size_of(b) > 3) { ++n_over; total += (size_t)size_of(b); }
    if (n_over) {
      cand_off.assign(nb, -1);
      size_t at = cand.size();
      for (size_t b = 0; b < nb; ++b) if (size_of(b) > 3) { cand_off[b] = (int32_t)at; at += (size_t)size_of(b); }
      cand.resize(cand.size() + total);
    }
  }
  const size_t ent0 = ent.size();
  ent.resize(ent0 + 4 * nb);
  parallel(nb, [&](size_t b_begin, size_t b_end) {
    int32_t* eo = ent.data() + ent0 + 4 * b_begin;
    for (size_t b = b_begin; b < b_end; ++b, eo += 4) {
      const int start = b ? cnt[b - 1] : 0, k = cnt[b] - start;
      int32_t* it = items.data() + start;
      if (k > 1) std::sort(it, it + k);
      eo[0] = -1; eo[1] = eo[2] = eo[3] = 0;
      if (k == 1 && sole[b]) { eo[0] = 0; eo[1] = it[0]; }
      else if (k >= 1 && k <= 3) { eo[0] = k; for (int t = 0; t < k; ++t) eo[1 + t] = it[t]; }
      else if (k > 3) { eo[0] = k; eo[1] = cand_off[b]; std::copy(it, it + k, cand.data() + cand_off[b]); }
    }
  });
  lap("entries");
  fs.ox = ox; fs.oy = oy; fs.inv_cx = inv_cx; fs.inv_cy = inv_cy; fs.nx = nx; fs.ny = ny;
  fs.bucket_off = (int32_t)(ent0 / 4); fs.poly_base = poly_base;
}

// One 96-byte record per polygon (rthx_kernels.cu: point_in_rec, wall_of_rec), read with three 32-byte loads:
//   [0..3] vx[4]   [4..7] vy[4]   (a triangle repeats vertex 0 in slot 3)
//   [8..11] {int32 vertex count, int32 bit i: the outward normal of edge i is -(ey, -ex) instead of (ey, -ex),
//            int32 surface id of walls 0..3 (cells; -1 otherwise), 8 bytes unused}
constexpr int FREC = 12;
inline void face_record(const Poly& p, const int32_t* surf4, double* rec) {
  for (int i = 0; i < 4; ++i) { const int k = i < p.n ? i : 0; rec[i] = p.vx[k]; rec[4 + i] = p.vy[k]; }
  int32_t tail[8] = {p.n, 0, -1, -1, -1, -1, 0, 0};
  for (int i = 0; i < p.n; ++i) {
    const int j = (i + 1) % p.n;
    const double ex = p.vx[j] - p.vx[i], ey = p.vy[j] - p.vy[i];
    if (ey * p.nx[i] - ex * p.ny[i] < 0) tail[1] |= 1 << i;
  }
  if (surf4) for (int i = 0; i < 4; ++i) tail[2 + i] = surf4[i];
  std::memcpy(rec + 8, tail, 32);
}

// ---- host replay of the device locator (rthx_kernels.cu: point_in_rec, find_face_generic, wall_of_rec), statement for statement;
// ---- used by tests/grid_check.cpp only
inline bool host_point_in_rec(const double* rec, double px, double py) {
  const double* vx = rec; const double* vy = rec + 4;
  unsigned inside = 0u;
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 3) & 3;
    const double xi = vx[i], yi = vy[i], xj = vx[j], yj = vy[j];
    const bool straddle = (yi > py) != (yj > py);
    const double dyv = yj - yi;
    const double t = std::fma(px - xi, dyv, -((xj - xi) * (py - yi)));
    const bool crossed = straddle && (t != 0.0) && (std::signbit(t) != std::signbit(dyv));
    inside ^= crossed ? 1u : 0u;
  }
  return inside != 0u;
}

inline int host_find_face(const FaceSetDev& fs, const int32_t* ent, const int32_t* cand, const double* face_rec, double px, double py,
                          int* code_out = nullptr) {
  const double fi = std::floor((px - fs.ox) * fs.inv_cx), fj = std::floor((py - fs.oy) * fs.inv_cy);
  if (code_out) *code_out = -2;
  if (!(fi >= 0.0 && fi < (double)fs.nx && fj >= 0.0 && fj < (double)fs.ny)) return -1;
  const int32_t* e = ent + 4 * ((size_t)fs.bucket_off + (size_t)fi + (size_t)fj * fs.nx);
  if (code_out) *code_out = e[0];
  if (e[0] == 0) return e[1];
  const double* recs = face_rec + (size_t)FREC * fs.poly_base;
  if (e[0] <= 3) {
    for (int k = 0; k < e[0]; ++k)
      if (host_point_in_rec(recs + (size_t)FREC * e[1 + k], px, py)) return e[1 + k];
    return -1;
  }
  for (int k = 0; k < e[0]; ++k) {
    const int f = cand[e[1] + k];
    if (host_point_in_rec(recs + (size_t)FREC * f, px, py)) return f;
  }
  return -1;
}

inline int host_wall_of_rec(const double* rec, double px, double py, double dx, double dy, int* surf) {
  const double* vx = rec; const double* vy = rec + 4;
  int32_t tail[8];
  std::memcpy(tail, rec + 8, 32);
  const int nv = tail[0], bits = tail[1];
  double bn = 1.0, bd = 0.0;
  int bi = 0;
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    const double ex = vx[j] - vx[i], ey = vy[j] - vy[i];
    const bool flip = (bits >> i) & 1;
    const double ax = flip ? -ey : ey, ay = flip ? ex : -ex;
    const double den = dx * ax + dy * ay;
    const double num = (vx[i] - px) * ax + (vy[i] - py) * ay;
    const double an = std::fabs(num), ad = std::fabs(den);
    const bool better = (i < nv) && (den * den >= 1e-20 * (ex * ex + ey * ey)) && (num * den > 0.0) && (an * bd < bn * ad);
    if (better) { bn = an; bd = ad; bi = i; }
  }
  *surf = tail[2 + bi];
  return bi;
}

}  // namespace rthx
