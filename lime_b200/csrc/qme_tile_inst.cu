// explicit instantiations of the register-patch cluster propagator (see qme_tile.cuh)
#include "qme_tile.cuh"

int qme_tile_launch(const QmeTileArgs& a, int NP, int S, size_t smem, cudaStream_t st) {
    if (NP == 128) {
        if (S == 0) return qme_tile_launch_one<128, 4, 0>(a, smem, st);
        if (S == 1) return qme_tile_launch_one<128, 4, 1>(a, smem, st);
        return qme_tile_launch_one<128, 4, 2>(a, smem, st);
    }
    if (S == 0) return qme_tile_launch_one<64, 4, 0>(a, smem, st);
    if (S == 1) return qme_tile_launch_one<64, 4, 1>(a, smem, st);
    return qme_tile_launch_one<64, 4, 2>(a, smem, st);
}
