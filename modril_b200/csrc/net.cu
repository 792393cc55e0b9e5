// frl_net lifetime + weight packing (flat fp32 master vector -> kernel layouts).
// Replaces what the reference leaves to nn.Module construction / load_state_dict
// (modril/flowril/networks.py:10-21, :34-44; modril/flowril/flowril.py:42-49).
#include <mutex>
#include <vector>

#include "frl_internal.h"

namespace frl {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_last_error = buf;
    return FRL_ERR_CUDA;
}

// One copy job: src matrix [rows][cols] (row-major, inside the flat parameter vector) -> dst, either same
// orientation (dst[(r + dst_row0) * dst_ld + c]) or transposed (dst[c * dst_ld + r + dst_row0]).
struct PackJob {
    long long src_off;
    long long dst_off;
    int rows, cols, dst_ld, dst_row0, transpose;
    long long first;  // prefix sum of element counts
};
constexpr int kMaxJobs = 40;
struct PackTable {
    PackJob job[kMaxJobs];
    int n_jobs;
    long long total;
};

__global__ void pack_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, PackTable tab) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tab.total;
         i += (long long)gridDim.x * blockDim.x) {
        int j = 0;
        while (j + 1 < tab.n_jobs && tab.job[j + 1].first <= i) ++j;
        const PackJob& jb = tab.job[j];
        long long e = i - jb.first;
        int r = (int)(e / jb.cols), c = (int)(e % jb.cols);
        float v = src[jb.src_off + e];
        if (jb.transpose)
            dst[jb.dst_off + (long long)c * jb.dst_ld + r + jb.dst_row0] = v;
        else
            dst[jb.dst_off + (long long)(r + jb.dst_row0) * jb.dst_ld + c] = v;
    }
}

}  // namespace frl

using namespace frl;

extern "C" int frl_abi_version(void) { return FRL_ABI_VERSION; }
extern "C" const char* frl_last_error(void) { return g_last_error.c_str(); }

extern "C" int frl_net_create(const frl_net_config* cfg, frl_net** out) {
    FRL_REQUIRE(cfg && out, "null config/out");
    FRL_REQUIRE(cfg->s_dim >= 1 && cfg->a_dim >= 1, "s_dim and a_dim must be >= 1");
    FRL_REQUIRE(cfg->a_dim <= 32, "a_dim > 32 unsupported");
    FRL_REQUIRE(cfg->s_dim + cfg->a_dim + 1 <= 256, "a_dim + s_dim + 1 > 256 unsupported");
    FRL_REQUIRE(cfg->hidden == 128 || cfg->hidden == 256, "hidden must be 128 or 256");
    FRL_REQUIRE(cfg->depth >= 1 && cfg->depth <= kMaxDepth, "depth must be in 1..8");
    FRL_REQUIRE(cfg->n_heads == 1 || cfg->n_heads == 2, "n_heads must be 1 or 2");
    int ndev = 0;
    FRL_CUDA(cudaGetDeviceCount(&ndev));
    FRL_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "bad device ordinal");
    FRL_CUDA(cudaSetDevice(cfg->device));

    frl_net* n = new (std::nothrow) frl_net();
    if (!n) { set_error("out of host memory"); return FRL_ERR_NOMEM; }
    n->cfg = *cfg;
    const int H = cfg->hidden, A = cfg->a_dim;
    n->d_in = cfg->a_dim + cfg->s_dim + 1;
    n->d_in_p = round_up(n->d_in, 16);
    n->n_out = cfg->n_heads * A;
    n->n_out_p = round_up(n->n_out, 16);

    long long off = 0;
    for (int l = 0; l < cfg->depth; ++l) {
        int K = l == 0 ? n->d_in : H;
        n->off_w[l] = off; off += (long long)H * K;
        n->off_b[l] = off; off += H;
    }
    for (int h = 0; h < cfg->n_heads; ++h) {
        n->off_wh[h] = off; off += (long long)A * H;
        n->off_bh[h] = off; off += A;
    }
    n->n_params = off;

    // FP32 pack: per layer Wt [Kp][H], W [H][Kp], b [H]; heads Wh [NOp][H], Wht [H][NOp], bh [NOp]
    size_t floats = 0;
    std::vector<size_t> o_wt(cfg->depth), o_w(cfg->depth), o_b(cfg->depth);
    for (int l = 0; l < cfg->depth; ++l) {
        int Kp = l == 0 ? n->d_in_p : H;
        o_wt[l] = floats; floats += (size_t)Kp * H;
        o_w[l] = floats; floats += (size_t)H * Kp;
        o_b[l] = floats; floats += H;
    }
    size_t o_wh = floats; floats += (size_t)n->n_out_p * H;
    size_t o_wht = floats; floats += (size_t)H * n->n_out_p;
    size_t o_bh = floats; floats += n->n_out_p;
    n->pack_f32_floats = floats;
    cudaError_t e = cudaMalloc(&n->pack_f32, floats * sizeof(float));
    if (e != cudaSuccess) { delete n; return cuda_fail(e, "cudaMalloc(pack_f32)", __FILE__, __LINE__); }
    e = cudaMemset(n->pack_f32, 0, floats * sizeof(float));
    if (e != cudaSuccess) { cudaFree(n->pack_f32); delete n; return cuda_fail(e, "cudaMemset", __FILE__, __LINE__); }
    for (int l = 0; l < cfg->depth; ++l) {
        n->f32.Wt[l] = n->pack_f32 + o_wt[l];
        n->f32.W[l] = n->pack_f32 + o_w[l];
        n->f32.b[l] = n->pack_f32 + o_b[l];
    }
    n->f32.Wh = n->pack_f32 + o_wh;
    n->f32.Wht = n->pack_f32 + o_wht;
    n->f32.bh = n->pack_f32 + o_bh;
    *out = n;
    return FRL_OK;
}

extern "C" int frl_net_destroy(frl_net* net) {
    if (!net) return FRL_OK;
    cudaSetDevice(net->cfg.device);
    if (net->pack_f32) cudaFree(net->pack_f32);
    if (net->pack_tc) cudaFree(net->pack_tc);
    if (net->pack_tc3) cudaFree(net->pack_tc3);
    if (net->pack_tc3x) cudaFree(net->pack_tc3x);
    if (net->ws) cudaFree(net->ws);
    if (net->pack_train) cudaFree(net->pack_train);
    for (int k = 0; k < PACK_KINDS; ++k)
        if (net->pack_event[k]) cudaEventDestroy(net->pack_event[k]);
    if (net->ws_tc) cudaFree(net->ws_tc);
    if (net->ws_x0s) cudaFree(net->ws_x0s);
    for (void* r : net->retired) cudaFree(r);
    for (int i = 0; i < frl_net::kMaxReaders; ++i)
        if (net->reader_event[i]) cudaEventDestroy(net->reader_event[i]);
    if (net->x0s_event) cudaEventDestroy(net->x0s_event);
    frl_host_stage_free(net);
    delete net;
    return FRL_OK;
}

extern "C" long long frl_net_num_params(const frl_net* net) { return net ? net->n_params : -1; }

namespace frl {

int pack_f32(frl_net* net, const float* params_dev, cudaStream_t st) {
    const int H = net->cfg.hidden, A = net->cfg.a_dim;
    PackTable tab;
    int nj = 0;
    long long total = 0;
    auto add = [&](long long src_off, const float* dst, int rows, int cols, int dst_ld, int dst_row0, int transpose) {
        PackJob& j = tab.job[nj++];
        j.src_off = src_off;
        j.dst_off = dst - net->pack_f32;
        j.rows = rows; j.cols = cols; j.dst_ld = dst_ld; j.dst_row0 = dst_row0; j.transpose = transpose;
        j.first = total;
        total += (long long)rows * cols;
    };
    for (int l = 0; l < net->cfg.depth; ++l) {
        int K = l == 0 ? net->d_in : H;
        int Kp = l == 0 ? net->d_in_p : H;
        add(net->off_w[l], net->f32.Wt[l], H, K, H, 0, 1);    // Wt[k][n] = W[n][k]
        add(net->off_w[l], net->f32.W[l], H, K, Kp, 0, 0);    // W[n][k], padded leading dim
        add(net->off_b[l], net->f32.b[l], 1, H, H, 0, 0);
    }
    for (int h = 0; h < net->cfg.n_heads; ++h) {
        add(net->off_wh[h], net->f32.Wh, A, H, H, h * A, 0);                 // Wh[o][k]
        add(net->off_wh[h], net->f32.Wht, A, H, net->n_out_p, h * A, 1);     // Wht[k][o]
        add(net->off_bh[h], net->f32.bh, 1, A, net->n_out_p, 0, 0);
        tab.job[nj - 1].dst_off += h * A;                                     // bias h lands at bh[h*A ...]
    }
    tab.n_jobs = nj;
    tab.total = total;
    int threads = 256;
    int blocks = (int)((total + threads - 1) / threads);
    if (blocks > 592) blocks = 592;
    pack_f32_kernel<<<blocks, threads, 0, st>>>(params_dev, net->pack_f32, tab);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

static int is_capturing(cudaStream_t st, bool* out) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    FRL_CUDA(cudaStreamIsCapturing(st, &cap));
    *out = cap != cudaStreamCaptureStatusNone;
    return FRL_OK;
}

// Record that kernels reading this net's packed layouts have just been launched on `st`.  (Inside a stream capture the
// graph's own edges order a later re-pack captured in the same graph; events cannot cross the capture boundary.)
int note_reader(const frl_net* cnet, cudaStream_t st) {
    frl_net* net = const_cast<frl_net*>(cnet);
    bool cap = false;
    int rc = is_capturing(st, &cap);
    if (rc != FRL_OK || cap) return rc;
    int slot = -1;
    for (int i = 0; i < frl_net::kMaxReaders; ++i)
        if (net->reader_used[i] && net->reader_stream[i] == st) slot = i;
    if (slot < 0)
        for (int i = 0; i < frl_net::kMaxReaders && slot < 0; ++i)
            if (!net->reader_used[i]) slot = i;
    if (slot < 0) {   // more distinct reader streams than slots: retire the oldest one by waiting for it here
        slot = 0;
        FRL_CUDA(cudaEventSynchronize(net->reader_event[0]));
    }
    if (!net->reader_event[slot]) FRL_CUDA(cudaEventCreateWithFlags(&net->reader_event[slot], cudaEventDisableTiming));
    FRL_CUDA(cudaEventRecord(net->reader_event[slot], st));
    net->reader_stream[slot] = st;
    net->reader_used[slot] = true;
    return FRL_OK;
}

// Build the layout `kind` from the current master weights if it is stale.  Runs on the consumer's stream, so the pack
// is ordered before the kernel that needs it; a different stream that finds the layout up to date waits on the event
// recorded after the pack, and the pack itself first waits for the readers of the old image on other streams.
int ensure_pack(const frl_net* cnet, int kind, cudaStream_t st) {
    frl_net* net = const_cast<frl_net*>(cnet);
    bool cap = false;
    int rc = is_capturing(st, &cap);
    if (rc != FRL_OK) return rc;
    if (net->packed[kind] == net->version) {
        if (!cap && net->pack_stream[kind] != st && net->pack_event[kind])
            FRL_CUDA(cudaStreamWaitEvent(st, net->pack_event[kind], 0));
        return FRL_OK;
    }
    if (!cap)
        for (int i = 0; i < frl_net::kMaxReaders; ++i)
            if (net->reader_used[i] && net->reader_stream[i] != st) FRL_CUDA(cudaStreamWaitEvent(st, net->reader_event[i], 0));
    switch (kind) {
        case PACK_F32: rc = pack_f32(net, net->params_dev, st); break;
        case PACK_TC1: rc = tc_pack(net, net->params_dev, st); break;
        case PACK_TC2: set_error("pack kind 2 is retired"); return FRL_ERR_INVALID;
        case PACK_TC3: rc = tc3_pack(net, net->params_dev, st); break;
        case PACK_TC3X: rc = tc3x_pack(net, net->params_dev, st); break;
        case PACK_TRAIN: rc = net->n_out_p <= 64 ? tc_train_pack(net, net->params_dev, st) : FRL_OK; break;
        default: set_error("bad pack kind"); return FRL_ERR_INVALID;
    }
    if (rc != FRL_OK) return rc;
    if (!cap) {   // an event recorded inside a capture could not be waited on from outside of it
        if (!net->pack_event[kind]) FRL_CUDA(cudaEventCreateWithFlags(&net->pack_event[kind], cudaEventDisableTiming));
        FRL_CUDA(cudaEventRecord(net->pack_event[kind], st));
    }
    net->pack_stream[kind] = st;
    net->packed[kind] = net->version;
    return FRL_OK;
}

}  // namespace frl

extern "C" int frl_net_set_params(frl_net* net, const float* params_dev, void* /*stream*/) {
    FRL_REQUIRE(net && params_dev, "null net/params");
    net->params_dev = params_dev;
    net->version += 1;     // every kernel family re-packs its own layout on next use (frl::ensure_pack)
    net->has_params = true;
    return FRL_OK;
}
