// step.cu -- vehicle-step kernel (placeholder until the mj_step restatement lands).
#include "common.h"
using namespace ftgp;
extern "C" int ftgp_step(const ftgp_geom*, double*, double*, double*, const double*, const int32_t*, int64_t, int,
                         int32_t*, void*) {
    set_error("ftgp_step: not built yet"); return FTGP_ERR_UNSUPPORTED;
}
extern "C" int ftgp_tick(const ftgp_tick_args*, int, void*) {
    set_error("ftgp_tick: not built yet"); return FTGP_ERR_UNSUPPORTED;
}
