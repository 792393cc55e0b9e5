// tcgen05 tensor-core log-density integrator (placeholder until the UMMA kernel lands).
#include "frl_internal.h"

namespace frl {

int tc_pack(frl_net* net, const float* params_dev, cudaStream_t stream) {
    (void)net; (void)params_dev; (void)stream;
    return FRL_OK;
}

int launch_logprob_tc(const ScoreArgs& args, int precision, cudaStream_t stream) {
    (void)args; (void)precision; (void)stream;
    set_error("tensor-core precision modes are not built into this library yet");
    return FRL_ERR_UNSUPPORTED;
}

}  // namespace frl
