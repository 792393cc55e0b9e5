// Rademacher probes of a whole rollout in ONE launch, bit-identical to the sequence of small draws the reference makes.
//
// Reference: FlowMatching._hutch_div (modril/flowril/pretrain.py:88-92) draws  v = randint_like(x, 0, 2) * 2 - 1  once per
// Euler step and per net, and the reward hook calls log_prob once per rollout step (flowril.py:186-189, base_irl.py:53-55):
// T rollout steps x 2 nets x n_steps draws of [N, a] values from torch's CUDA generator = 8192 tiny launches for the
// default rollout.  torch's generator is Philox4x32-10 keyed by (seed, element index, offset), and a draw of
// numel <= 256 * (resident blocks) values advances the offset by 4 and gives element i the first 32-bit word of
// curand4(curand_init(seed, i, offset)), reduced modulo the range (ATen/native/cuda/DistributionTemplates.h:
// distribution_nullary_kernel with unroll 4, random_from_to).  So draw number c of the sequence is a pure function of
// (seed, offset0 + 4 c, i): this kernel evaluates all of them at once and writes them in the [n_steps][T * N][a] layout
// the integrators take (FRL_PROBE_PER_STEP).  The host side (modril_b200/flowril/flowril.py) reads seed / offset from the
// generator, advances the offset by 4 * (number of draws) afterwards, and verifies the replay against real randint_like
// calls once per process (falling back to them if torch ever changes its kernel).
#include <curand_kernel.h>

#include <algorithm>

#include "frl_internal.h"

namespace frl {

// call index c = (outer * n_nets + net) * n_steps + k;  out[net][k][outer][i]
__global__ void rademacher_calls_kernel(unsigned long long seed, unsigned long long offset0, int n_outer, int n_nets, int n_steps,
                                        int numel, float* __restrict__ out0, float* __restrict__ out1) {
    const long long total = (long long)n_outer * n_nets * n_steps * numel;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % numel);
        const long long c = e / numel;
        const int k = (int)(c % n_steps);
        const int net = (int)((c / n_steps) % n_nets);
        const int outer = (int)(c / ((long long)n_steps * n_nets));
        curandStatePhilox4_32_10_t state;
        curand_init(seed, (unsigned long long)i, offset0 + 4ull * (unsigned long long)c, &state);
        const uint4 r = curand4(&state);
        float* out = net == 0 ? out0 : out1;
        out[((long long)k * n_outer + outer) * numel + i] = (r.x & 1u) ? 1.f : -1.f;
    }
}

}  // namespace frl

extern "C" int frl_rademacher_calls(unsigned long long seed, unsigned long long offset, int n_outer, int n_nets, int n_steps,
                                    int numel, float* out0_dev, float* out1_dev, int device, void* stream) {
    using namespace frl;
    FRL_REQUIRE(n_outer >= 1 && n_steps >= 1 && numel >= 1, "empty draw");
    FRL_REQUIRE(n_nets == 1 || n_nets == 2, "one or two nets");
    FRL_REQUIRE(out0_dev != nullptr && (n_nets == 1 || out1_dev != nullptr), "null output");
    // one element per generator thread in the reference's launch: numel <= 256 * (blocks torch would launch); 65536
    // values per draw (2048 environments x 32 action dims) is far inside that on any GPU this library runs on
    FRL_REQUIRE(numel <= 65536, "draws of more than 65536 values are not replayed");
    FRL_CUDA(cudaSetDevice(device));
    const long long total = (long long)n_outer * n_nets * n_steps * numel;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148 * 8);
    rademacher_calls_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(seed, offset, n_outer, n_nets, n_steps, numel, out0_dev,
                                                                        out1_dev);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

extern "C" int frl_reward(const float* logp_e_dev, const float* logp_a_dev, float* reward_dev, long long B, int reward_type,
                          int device, void* stream) {
    using namespace frl;
    FRL_REQUIRE(logp_e_dev && logp_a_dev && reward_dev, "null argument");
    FRL_REQUIRE(reward_type >= 0 && reward_type <= 4, "unknown reward type");
    if (B <= 0) return FRL_OK;
    FRL_CUDA(cudaSetDevice(device));
    return launch_reward(logp_e_dev, logp_a_dev, reward_dev, B, reward_type, (cudaStream_t)stream);
}
