// explicit instantiations of the chain-structured register-tiled cluster propagator (see qme_band.cuh)
#include "qme_band.cuh"
QME_BAND_DEFINE_LAUNCH_CHAIN(qme_band_launch_chain_tc2, 2)
