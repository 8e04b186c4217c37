// gp_args.h -- launch arguments of gp_tile_kernel (gp_tile.cuh), shared between the per-dimension
// translation units (gp_tile_inst.cu) and the dispatcher (gp_sweep.cu).
#pragma once
#include <stdint.h>

enum { MODE_SWEEP_GRID = 0, MODE_SWEEP_STATES = 1, MODE_PREDICT = 2 };

struct slb_gp_args {
    const double* points;   // MODE_SWEEP_STATES: [n, d]; MODE_PREDICT: [n, d_in]
    int64_t n;
    int64_t idx_begin;
    int32_t mode;
    int32_t want_var;
    uint8_t* negative;
    double* values;
    double* decrease;
    double* threshold;
    double* mean;
    double* err;
    const int64_t* index_list;           // refine mode: the tile's points are index_list[rel]
    const unsigned long long* count;     // refine mode: number of list entries (read on the device)
    int64_t count_min, count_max;        // refine mode: this launch works iff count_min < *count <= count_max
    long long* timing;      // diagnostics: [tile][warp][8]: cycles in {generate, contract, epilogue, total}, globaltimer ns {start, end}, cycles waiting at barriers, 0
};


