// gp_tile_inst.cu -- one translation unit per GP input dimension (compiled with
// -DSLB_TILE_DIN=1..6, in parallel): the instantiations of gp_tile_kernel (gp_tile.cuh) for that
// dimension -- plain RBF, covariance expressions, and (d_in = 3) the phase-timing build.
#include "gp_tile.cuh"

#ifndef SLB_TILE_DIN
#error "compile with -DSLB_TILE_DIN=<1..6>"
#endif

#define SLB_CAT2(a, b) a##b
#define SLB_CAT(a, b) SLB_CAT2(a, b)

int SLB_CAT(slb_gp_tile_launch_, SLB_TILE_DIN)(cudaStream_t st, const slb_sweep& cfg,
                                               const slb_gp_args& a, bool kexpr, bool timing) {
#if SLB_TILE_DIN == 3
    if (timing) return launch_gp_tile<3, true, false>(st, cfg, a);
#else
    if (timing) {
        slb_set_error("phase timing is compiled for d_in = 3 only");
        return 1;
    }
#endif
    return kexpr ? launch_gp_tile<SLB_TILE_DIN, false, true>(st, cfg, a)
                 : launch_gp_tile<SLB_TILE_DIN, false, false>(st, cfg, a);
}
