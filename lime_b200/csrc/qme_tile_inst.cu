// explicit instantiations of the register-patch cluster propagator (see qme_tile.cuh)
#include "qme_tile.cuh"
#include <cstdlib>

// kernel variant (qme_tile.cuh): bit 0 = clobber-free tensor-memory statements + early sandwich loads,
// bit 1 = warp-level stage synchronisation instead of the CTA barrier, bit 2 = row coefficients from shared memory and
// 16-word tensor-memory reads in stages 1-3 only (bit 0 is implied), bit 3 (with bit 2; value 8 = 12) = the thread's own rows of the stage
// vector from tensor memory too.  Built: 0, 1, 2, 3, 4, 6, 8.  LIMEB200_TILE_V overrides the default.
#ifndef QME_TILE_DEFAULT_V
#define QME_TILE_DEFAULT_V 0
#endif
int qme_tile_variant() {
    const char* e = getenv("LIMEB200_TILE_V");
    int v = QME_TILE_DEFAULT_V;
    if (e && *e >= '0' && *e <= '8' && !e[1]) v = *e - '0';
    if (v == 5) v = 4;
    if (v == 7) v = 6;
    return v;
}

template <int NP, int V>
static int launch_s(const QmeTileArgs& a, int S, size_t smem, cudaStream_t st) {
    if (S == 0) return qme_tile_launch_one<NP, 4, 0, V>(a, smem, st);
    if (S == 1) return qme_tile_launch_one<NP, 4, 1, V>(a, smem, st);
    return qme_tile_launch_one<NP, 4, 2, V>(a, smem, st);
}
template <int NP>
static int launch_v(const QmeTileArgs& a, int S, size_t smem, cudaStream_t st) {
    switch (qme_tile_variant()) {
        case 1: return launch_s<NP, 1>(a, S, smem, st);
        case 2: return launch_s<NP, 2>(a, S, smem, st);
        case 3: return launch_s<NP, 3>(a, S, smem, st);
        case 4: return launch_s<NP, 4>(a, S, smem, st);
        case 6: return launch_s<NP, 6>(a, S, smem, st);
        case 8: return launch_s<NP, 12>(a, S, smem, st);
        default: return launch_s<NP, 0>(a, S, smem, st);
    }
}

int qme_tile_launch(const QmeTileArgs& a, int NP, int S, size_t smem, cudaStream_t st) {
    if (NP == 128) return launch_v<128>(a, S, smem, st);
    return launch_v<64>(a, S, smem, st);
}
