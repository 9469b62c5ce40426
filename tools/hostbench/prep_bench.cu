// Host-only timing of rthx_create's mesh preparation (prepare_mesh in csrc/rthx_api.cu) — runs without a GPU.
// Build + run: python tools/hostbench/prep_bench.py
#define RTHX_PREP_BENCH 1
#include "../../raytraceheattransfer.jl_b200/csrc/rthx_api.cu"

extern "C" double rthx_prep_bench(const rthx_mesh* m, int reps, int generic) {
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    const auto t0 = std::chrono::steady_clock::now();
    std::shared_ptr<HostImage> im;
    std::string err;
    if (prepare_mesh(m, im, err) != RTHX_OK) { std::fprintf(stderr, "prepare_mesh: %s\n", err.c_str()); return -1.0; }
    (void)generic;
    const auto t1 = std::chrono::steady_clock::now();
    best = std::min(best, std::chrono::duration<double, std::milli>(t1 - t0).count());
  }
  return best;
}
