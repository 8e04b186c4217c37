// gp_args.h -- launch arguments of gp_tile_kernel (gp_tile.cuh), shared between the per-dimension
// translation units (gp_tile_inst.cu) and the dispatcher (gp_sweep.cu).
#pragma once
#include <stdint.h>

enum { MODE_SWEEP_GRID = 0, MODE_SWEEP_STATES = 1, MODE_PREDICT = 2 };
enum { SLB_SPLIT_ITEMS = 512, SLB_SPLIT_MAX = 8, SLB_SPLIT_TICKET_BYTES = 2048 };
// bytes of the split workspace: partial sums of SLB_SPLIT_ITEMS CTAs (64-point upper bound per tile)
#define SLB_SPLIT_PARTIAL_BYTES ((size_t)SLB_SPLIT_ITEMS * SLB_MAX_OUT * (1 + SLB_MAX_OUT) * 64 * sizeof(double))

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
    // refine mode, short lists: the rows of L^-1 of one point tile are split over up to `split_max`
    // CTAs (equal triangular areas); every CTA leaves its partial sums in `split_partial`
    // [CTA][factor][1 + MAX_OUT][tile points], the last one to arrive at `split_ticket[tile]` adds
    // them in group order and finishes the tile.  NULL / 0: no split.  Grids of at most
    // SLB_SPLIT_ITEMS CTAs.
    double* split_partial;
    int* split_ticket;
    int32_t split_max;
    int32_t split_factors;  // 1: with CTAs to spare, the factors of a tile go to separate CTAs before its rows are split
    long long* timing;      // diagnostics: [tile][warp][8]: cycles in {generate, contract, epilogue, total}, globaltimer ns {start, end}, cycles waiting at barriers, 0
};


