// explicit instantiations of the register-patch cluster propagator (see qme_tile.cuh)
#include "qme_tile.cuh"
#include <cstdlib>

// LIMEB200_TILE_V=8 selects the tensor-memory-window variant (the thread's own stage-vector rows from tensor memory,
// row coefficients from shared memory: 27 % fewer shared-memory wavefronts, measured +0.8 % -- within the noise, so
// the default stays the original form); anything else = default.
int qme_tile_variant() {
    const char* e = getenv("LIMEB200_TILE_V");
    return (e && e[0] == '8' && !e[1]) ? 8 : 0;
}

template <int NP, bool TW>
static int launch_s(const QmeTileArgs& a, int S, size_t smem, cudaStream_t st) {
    if (S == 0) return qme_tile_launch_one<NP, 4, 0, TW>(a, smem, st);
    if (S == 1) return qme_tile_launch_one<NP, 4, 1, TW>(a, smem, st);
    return qme_tile_launch_one<NP, 4, 2, TW>(a, smem, st);
}

int qme_tile_launch(const QmeTileArgs& a, int NP, int S, size_t smem, cudaStream_t st) {
    const bool tw = qme_tile_variant() == 8;
    if (NP == 128) return tw ? launch_s<128, true>(a, S, smem, st) : launch_s<128, false>(a, S, smem, st);
    return tw ? launch_s<64, true>(a, S, smem, st) : launch_s<64, false>(a, S, smem, st);
}
