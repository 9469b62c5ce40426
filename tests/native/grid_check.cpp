// Host check of the generic locator's tables (csrc/rthx_grid.h): for every face set of a mesh (coarse faces, and the fine cells of
// every coarse face) the bucket grid is built exactly as rthx_api.cu builds it, and the device's locator — replayed on the CPU
// statement for statement — is compared with a brute-force crossing-number scan over ALL faces of the set in ascending order
// (findFace2D.jl's result for a grid of one bucket) on random points, points near vertices and points on bucket boundaries.
// Also checks the wall lookup from vertices against the faithful form with unit normals.
// Test infrastructure: built with g++ by tests/test_generic_grid.py, no GPU needed.
#include <cstdio>
#include <random>

#include "rthx_grid.h"

using namespace rthx;

extern "C" int rthx_grid_check(const rthx_mesh* m, int points_per_set, uint64_t seed, double* stats /* [8] */) {
  const int nc = m->n_coarse, ncell = m->n_cells;
  std::vector<Poly> polys((size_t)ncell + nc);
  for (int g = 0; g < ncell; ++g) {
    Poly& p = polys[g];
    p.n = m->cell_nv[g];
    for (int i = 0; i < p.n; ++i) { p.vx[i] = m->cell_vx[4 * (size_t)g + i]; p.vy[i] = m->cell_vy[4 * (size_t)g + i]; }
    p.midx = m->cell_mid[2 * (size_t)g]; p.midy = m->cell_mid[2 * (size_t)g + 1]; p.volume = m->cell_volume[g];
    poly_finish(p);
  }
  for (int c = 0; c < nc; ++c) {
    Poly& p = polys[(size_t)ncell + c];
    p.n = m->coarse_nv[c];
    double sx = 0, sy = 0;
    for (int i = 0; i < p.n; ++i) { p.vx[i] = m->coarse_vx[4 * (size_t)c + i]; p.vy[i] = m->coarse_vy[4 * (size_t)c + i]; sx += p.vx[i]; sy += p.vy[i]; }
    p.midx = sx / p.n; p.midy = sy / p.n;
    double a = 0;
    for (int i = 0; i < p.n; ++i) { const int j = (i + 1) % p.n; a += p.vx[i] * p.vy[j] - p.vx[j] * p.vy[i]; }
    p.volume = 0.5 * std::fabs(a);
    poly_finish(p);
  }
  std::vector<FaceSetDev> sets(1 + (size_t)nc);
  std::vector<int32_t> ent, cand;
  build_grid(&polys[ncell], nc, ncell, sets[0], ent, cand);
  for (int c = 0; c < nc; ++c) build_grid(&polys[m->fine_off[c]], m->fine_off[c + 1] - m->fine_off[c], m->fine_off[c], sets[1 + c], ent, cand);
  std::vector<double> frec(polys.size() * FREC);
  for (size_t i = 0; i < polys.size(); ++i) face_record(polys[i], i < (size_t)ncell ? m->cell_surf_id + 4 * i : nullptr, &frec[FREC * i]);

  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  long long n_pts = 0, n_bad = 0, n_sole = 0, n_cand_tests = 0, n_found = 0, n_wall = 0, n_wall_bad = 0;
  size_t n_buckets = 0;
  for (size_t si = 0; si < sets.size(); ++si) {
    const FaceSetDev& fs = sets[si];
    const int nf = si == 0 ? nc : m->fine_off[si] - m->fine_off[si - 1];
    const Poly* faces = &polys[fs.poly_base];
    n_buckets += (size_t)fs.nx * fs.ny;
    const double wx = fs.nx / fs.inv_cx, wy = fs.ny / fs.inv_cy;
    for (int t = 0; t < points_per_set; ++t) {
      double px, py;
      const int mode = t % 8;
      if (mode == 6 && nf > 0) {            // next to a vertex of a random face
        const Poly& q = faces[(size_t)(U(rng) * nf) % nf];
        const int v = (int)(U(rng) * q.n) % q.n;
        px = q.vx[v] + (U(rng) - 0.5) * 1e-7 * wx; py = q.vy[v] + (U(rng) - 0.5) * 1e-7 * wy;
      } else if (mode == 7) {               // next to a bucket corner
        px = fs.ox + std::floor(U(rng) * fs.nx) / fs.inv_cx + (U(rng) - 0.5) * 1e-12 * wx;
        py = fs.oy + std::floor(U(rng) * fs.ny) / fs.inv_cy + (U(rng) - 0.5) * 1e-12 * wy;
      } else {
        px = fs.ox + U(rng) * wx; py = fs.oy + U(rng) * wy;
      }
      int code;
      const int got = host_find_face(fs, ent.data(), cand.data(), frec.data(), px, py, &code);
      int want = -1;
      for (int f = 0; f < nf; ++f)
        if (host_point_in_rec(&frec[FREC * ((size_t)fs.poly_base + f)], px, py)) { want = f; break; }
      ++n_pts;
      if (code == 0) ++n_sole;
      if (code > 0) n_cand_tests += code;
      if (want >= 0) ++n_found;
      if (got != want) {
        if (++n_bad <= 5) std::fprintf(stderr, "grid_check: set %zu point (%.17g, %.17g): grid says %d, scan says %d (code %d)\n", si, px, py, got, want, code);
      }
      // wall lookup of the containing cell for a random direction: vertices-only form against unit normals
      if (si > 0 && want >= 0) {
        const double ang = 6.283185307179586 * U(rng), dx = std::cos(ang), dy = std::sin(ang);
        const Poly& q = faces[want];
        double best = INFINITY; int bw = 0;
        bool margin_ok = true;
        for (int i = 0; i < q.n; ++i) {
          const double den = dx * q.nx[i] + dy * q.ny[i], num = (q.vx[i] - px) * q.nx[i] + (q.vy[i] - py) * q.ny[i];
          if (std::fabs(den) < 1e-10) { if (std::fabs(den) > 0.5e-10) margin_ok = false; continue; }
          const double u = num / den;
          if (u > 0 && u < best) { if (best - u < 1e-9 * std::fabs(u)) margin_ok = false; best = u; bw = i; }
          else if (u > 0 && u - best < 1e-9 * std::fabs(u)) margin_ok = false;
        }
        int surf;
        const int w = host_wall_of_rec(&frec[FREC * ((size_t)fs.poly_base + want)], px, py, dx, dy, &surf);
        if (margin_ok && std::isfinite(best)) {
          ++n_wall;
          if (w != bw || surf != m->cell_surf_id[4 * ((size_t)fs.poly_base + want) + bw]) {
            if (++n_wall_bad <= 5) std::fprintf(stderr, "grid_check: wall of cell %d: record says %d, normals say %d\n", fs.poly_base + want, w, bw);
          }
        }
      }
    }
  }
  stats[0] = (double)n_pts; stats[1] = (double)n_bad; stats[2] = (double)n_sole; stats[3] = (double)n_cand_tests;
  stats[4] = (double)n_found; stats[5] = (double)n_buckets; stats[6] = (double)n_wall; stats[7] = (double)n_wall_bad;
  return 0;
}
