// explicit instantiations of the register-tiled cluster propagator (see qme_band.cuh)
#include "qme_band.cuh"
QME_BAND_DEFINE_LAUNCH(qme_band_launch_tc4_n4, 4, 4)
