/* rthx_oracle.h — TEST INFRASTRUCTURE (see rthx_oracle.c). Shares the mesh/argument structs of include/rthx.h. */
#ifndef RTHX_ORACLE_H
#define RTHX_ORACLE_H
#include "rthx.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rthx_oracle_stats {
  uint64_t n_surface_gas, n_surface_wall, n_volume_gas, n_volume_wall; /* tallied rays by (emitter kind, ending) */
  uint64_t n_crossings;                                                /* coarse-face crossings */
  uint64_t n_lost;
  int32_t n_threads;
  int32_t pad_;
  double loop_seconds;                                                 /* wall time of the emitter loop alone (no mesh build, no zeroing) */
} rthx_oracle_stats;

/* 1: "reference-faithful" timing mode — per-thread xoshiro256++ instead of Philox, Dict-like row tally (rthx_oracle.c) */
void rthx_oracle_set_faithful(int on);
int rthx_oracle_get_faithful(void);

int rthx_oracle_trace(const rthx_mesh* m, const rthx_trace_args* a, uint64_t* counts, uint64_t* lost,
                      rthx_rec_out* rec, int n_threads, rthx_oracle_stats* st);
int rthx_oracle_shoot(const rthx_mesh* m, const rthx_trace_args* a, int emitter, int band, uint64_t ray_id, double out[6]);
int rthx_oracle_emit(const rthx_mesh* m, const rthx_trace_args* a, int emitter, int band, int64_t n, double* out);
int rthx_oracle_find_face(const rthx_mesh* m, int set, double px, double py);
double rthx_oracle_dist_to_surface(int n, const double* vx, const double* vy, double px, double py, double dx, double dy, int* idx);
void rthx_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double rthx_oracle_u52(uint32_t lo, uint32_t hi);
float rthx_oracle_u23(uint32_t w);
double rthx_oracle_u32(uint32_t w);

#ifdef __cplusplus
}
#endif
#endif
