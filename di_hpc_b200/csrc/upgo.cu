// upgo.cu -- UPGO loss, forward and backward, for sm_100a.
//
// Semantics: hpc_rll/origin/upgo.py:21-38 (upgo_returns), 40-70 (upgo_loss), 7-18 (tb_cross_entropy):
//     l_t   = [r_{t+1} + v_{t+2} >= v_{t+1}]  (t < T-1)
//     ret_{T-1} = r_{T-1} + v_T;   ret_t = r_t + l_t*ret_{t+1} + (1-l_t)*v_{t+1}
//     adv = rho*(ret - v[:-1])  (no_grad);   loss = -mean(adv * log_softmax(logits)[a])
//     dloss/dlogits = -(adv/(T*B)) * (onehot(a) - softmax)
// Replaces UpgoForward/UpgoBackward (src/rl_utils/upgo.cu:8-69) and the 4 kernels of
// include/hpc/rll/cuda/rl_utils/upgo_kernel.h:11-108 (block-per-row cross entropy writing a (T,B,N)
// grad buffer that the backward re-reads).
//
// Forward = upgo_rows_fwd (streaming log-softmax gather -> metric (T,B)) + upgo_scan (ScanPipe over
// value/reward/rho/metric with the {0,1} coefficient carried in registers; emits coef = -adv/n and the
// loss partials) + finaliser.  Backward = shared softmax-gradient row kernel (recompute from logits).
#include "scan_lookback.cuh"
#include "scan_pipe.cuh"
#include "softmax_rows.cuh"

namespace hpcrll {

// G1: one lane per row known at compile time (G == 1, N == KMAX * WIDTH): see ppo_rows_fwd
template <int KMAX, int WIDTH, bool G1 = false>
__global__ void __launch_bounds__(256) upgo_rows_fwd(const float* __restrict__ logits,
                                                      const int64_t* __restrict__ action,
                                                      float* __restrict__ metric, int64_t R, int N_, int G_,
                                                      int log2G_) {
    using Row = RowRegs<KMAX, WIDTH>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = G1 ? KMAX * WIDTH : N_, G = G1 ? 1 : G_, log2G = G1 ? 0 : log2G_;
    const int lig = G1 ? 0 : (lane & (G - 1)), gw = lane >> log2G;
    const int rows_per_warp = 32 >> log2G;
    const int rows_per_block = rows_per_warp * 8;
    constexpr bool PF = Row::NE <= 8;  // software pipeline (see softmax_rows.cu)
    Row rr, nx;
    int a, na = -1;
    // row index and row pointer advance by increments (no 64-bit multiply per row)
    const int64_t stride = static_cast<int64_t>(gridDim.x) * rows_per_block;
    int64_t base = static_cast<int64_t>(blockIdx.x) * rows_per_block;
    int64_t row = base + warp * rows_per_warp + gw;
    const float* pl = logits + row * N;
    const int64_t pstep = stride * N;
    rr.load(pl, N, G, lig, row < R);
    a = row < R ? static_cast<int>(action[row]) : -1;
    for (; base < R; base += stride, row += stride) {  // block-uniform trip count
        const bool active = row < R;
        const int64_t nrow = row + stride;
        pl += pstep;
        if (PF) {
            nx.load(pl, N, G, lig, nrow < R);
            na = nrow < R ? static_cast<int>(action[nrow]) : -1;
        }
        const float m = rr.row_max(G);
        float s, t, none[Row::NE];
        rr.template stats<false, false>(G, m, s, t, none);
        const float sel = row_logp<false>(group_sum(rr.select(a, G, lig), G), m, logf(s));
        if (active && lig == 0) metric[row] = sel;
        if (PF) {
            rr = nx;
            a = na;
        } else {
            rr.load(pl, N, G, lig, nrow < R);
            a = nrow < R ? static_cast<int>(action[nrow]) : -1;
        }
    }
}

// staged variant (N <= 32, not 128-bit eligible): see softmax_rows.cuh "Staged rows"
__global__ void __launch_bounds__(kStageRows) upgo_rows_fwd_staged(const float* __restrict__ logits,
                                                                    const int64_t* __restrict__ action,
                                                                    float* __restrict__ metric, int64_t R, int N, int P,
                                                                    int aligned) {
    extern __shared__ float tile[];
    const int64_t ntiles = (R + kStageRows - 1) / kStageRows;
    for (int64_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
        const int64_t row0 = tix * kStageRows, row = row0 + threadIdx.x;
        __syncthreads();
        stage_rows(logits, R, N, P, row0, tile, aligned != 0);
        __syncthreads();
        if (row < R) {
            const float* x = tile + threadIdx.x * P;
            float m, s, t;
            staged_stats<false>(x, N, m, s, t);
            metric[row] = row_logp<false>(x[action[row]], m, logf(s));
        }
    }
}

__global__ void __launch_bounds__(256) upgo_rows_fwd_loop(const float* __restrict__ logits,
                                                           const int64_t* __restrict__ action,
                                                           float* __restrict__ metric, int64_t R, int N) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp; row < R; row += static_cast<int64_t>(gridDim.x) * 8) {
        const float* x = logits + row * N;
        float m = -INFINITY;
        for (int k = lane; k < N; k += 32) m = fmaxf(m, x[k]);
        m = warp_max(m);
        float s = 0.f;
        for (int k = lane; k < N; k += 32) s += exp_term(x[k] - m);
        s = warp_sum(s);
        if (lane == 0) metric[row] = row_logp<false>(x[action[row]], m, logf(s));
    }
}

struct UpgoBody {
    float ret, v1, v2, r1, neg_inv_n;
    float a_last;  // coefficient of the last step (0 or 1): used by the T-split look-back
    double acc;
    float* coef;  // running pointer, t descending
    int64_t ld;
    int t_last;
    bool valid;
    __device__ __forceinline__ void step(int t, const float (&x)[4], const float (&)[1]) {
        // x[0]=v_t x[1]=r_t x[2]=rho_t x[3]=metric_t
        const float v0 = x[0], r = x[1];
        if (t == t_last) {
            ret = __fadd_rn(r, v1);
            a_last = 0.f;
        } else {
            const float l = __fadd_rn(r1, v2) >= v1 ? 1.f : 0.f;
            a_last = l;
            ret = __fadd_rn(__fadd_rn(r, __fmul_rn(l, ret)), __fmul_rn(1.f - l, v1));
        }
        const float adv = __fmul_rn(x[2], __fsub_rn(ret, v0));
        acc += static_cast<double>(__fmul_rn(adv, x[3]));
        if (valid) st_stream(coef, adv * neg_inv_n);
        coef -= ld;
        v2 = v1;
        v1 = v0;
        r1 = r;
    }
};

template <int BT, int TT, int ST>
__global__ void __launch_bounds__(BT + 32) upgo_scan_tma(const __grid_constant__ TmapPack<4> maps,
                                                          const float* __restrict__ value, float* __restrict__ coef,
                                                          double* __restrict__ partials, int T, int B,
                                                          float neg_inv_n) {
    using Pipe = ScanPipe<4, BT, TT, ST, 0>;
    __shared__ double red[32];
    const int col0 = blockIdx.x * BT;
    const int col = col0 + threadIdx.x;
    UpgoBody body;
    body.valid = threadIdx.x < BT && col < B;
    body.ret = body.v2 = body.r1 = 0.f;
    body.neg_inv_n = neg_inv_n;
    body.acc = 0.0;
    body.ld = B;
    body.t_last = T - 1;
    body.coef = coef + static_cast<int64_t>(T - 1) * B + col;
    body.v1 = body.valid ? __ldg(value + static_cast<int64_t>(T) * B + col) : 0.f;
    Pipe::template run<true>(maps, nullptr, T, col0, body);
    double v[1] = {body.valid ? body.acc : 0.0};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = v[0];
}

// small batches: single-launch T-split with look-back (scan_lookback.cuh).  The coefficient is the 0/1 "keep following
// the trajectory" flag, so a segment publishes (prod a_t, zero-carry return); with a_t in {0,1} the composition is exact.
struct UpgoLbFac {
    const float* value;
    const float* reward;
    float* coef_out;
    int B, T, col;
    bool valid;
    float neg_inv_n;
    double* partial;
    using Body = UpgoBody;
    __device__ __forceinline__ Body make(int pass, int t_edge, float carry) const {
        Body b;
        b.valid = valid && pass == 1;
        b.ret = carry;
        b.neg_inv_n = neg_inv_n;
        b.a_last = 1.f;
        b.acc = 0.0;
        b.ld = B;
        b.t_last = T - 1;
        b.coef = coef_out + static_cast<int64_t>(t_edge - 1) * B + col;
        b.v1 = valid ? __ldg(value + static_cast<int64_t>(t_edge) * B + col) : 0.f;
        // history the first step of the segment looks at: r_{t+1}, v_{t+2} (rows that exist whenever t_edge < T)
        b.v2 = (valid && t_edge < T) ? __ldg(value + static_cast<int64_t>(t_edge + 1) * B + col) : 0.f;
        b.r1 = (valid && t_edge < T) ? __ldg(reward + static_cast<int64_t>(t_edge) * B + col) : 0.f;
        return b;
    }
    __device__ __forceinline__ void step(Body& b, int t, const float (&x)[4]) const {
        const float none[1] = {0.f};
        b.step(t, x, none);
    }
    static __device__ __forceinline__ float state(const Body& b) { return b.ret; }
    static __device__ __forceinline__ float coef(const Body& b) { return b.a_last; }
    __device__ __forceinline__ void finish(Body& b, int, int) const {
        double v = valid ? b.acc : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) *partial = v;
    }
};

__global__ void __launch_bounds__(kLbCols) upgo_scan_lookback(const float* __restrict__ value,
                                                              const float* __restrict__ reward,
                                                              const float* __restrict__ rho,
                                                              const float* __restrict__ metric, float* __restrict__ coef,
                                                              double* __restrict__ partials, int T, int B, float neg_inv_n,
                                                              int S, int L, int tiles, LbCtl* ctl,
                                                              unsigned long long* words) {
    __shared__ float smem[4 * kLbChunkRows * kLbCols];
    const LbTile lt = lb_begin(ctl, tiles);
    const int seg = S - 1 - lt.k;
    const int t0 = seg * L, t1 = min(T, t0 + L);
    const int col = lt.tile * kLbCols + threadIdx.x;
    const UpgoLbFac fac{value, reward, coef, B, T, col, col < B, neg_inv_n, partials + lt.vid};
    const float* const in[4] = {value, reward, rho, metric};
    const int64_t ld[4] = {B, B, B, B};
    lb_segment<4, true, false>(fac, in, ld, t0, t1, col, col < B, lt, words, tiles * kLbCols, 1.f, smem);
    lb_end(ctl, S * tiles, lt.epoch);
}

__global__ void __launch_bounds__(128) upgo_scan_generic(const float* __restrict__ value,
                                                          const float* __restrict__ reward,
                                                          const float* __restrict__ rho,
                                                          const float* __restrict__ metric, float* __restrict__ coef,
                                                          double* __restrict__ partials, int T, int B,
                                                          float neg_inv_n) {
    __shared__ double red[32];
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    UpgoBody body;
    body.valid = col < B;
    body.ret = body.v2 = body.r1 = 0.f;
    body.neg_inv_n = neg_inv_n;
    body.acc = 0.0;
    body.ld = B;
    body.t_last = T - 1;
    body.coef = coef + static_cast<int64_t>(T - 1) * B + col;
    if (body.valid) {
        body.v1 = value[static_cast<int64_t>(T) * B + col];
        const float none[1] = {0.f};
        for (int t = T - 1; t >= 0; --t) {
            const int64_t o = static_cast<int64_t>(t) * B + col;
            const float x[4] = {value[o], reward[o], rho[o], metric[o]};
            body.step(t, x, none);
        }
    }
    double v[1] = {body.valid ? body.acc : 0.0};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = v[0];
}

template <int BT, int TT, int ST>
static int launch_upgo_scan(const float* value, const float* reward, const float* rho, const float* metric,
                            float* coef, double* partials, int64_t T, int64_t B, float nin, cudaStream_t stream,
                            int* nblocks) {
    using Pipe = ScanPipe<4, BT, TT, ST, 0>;
    static SmemOptIn opt;
    auto kernel = upgo_scan_tma<BT, TT, ST>;
    if (int rc0 = opt.ensure(kernel, Pipe::kSmemBytes)) return rc0;
    TmapPack<4> maps;
    const float* srcs[4] = {value, reward, rho, metric};
    for (int k = 0; k < 4; ++k) {
        int rc1 = make_tmap_2d(&maps.m[k], srcs[k], k == 0 ? T + 1 : T, B, B, TT, BT);
        if (rc1) return rc1;
    }
    *nblocks = static_cast<int>((B + BT - 1) / BT);
    kernel<<<static_cast<unsigned>(*nblocks), Pipe::kThreads, Pipe::kSmemBytes, stream>>>(
        maps, value, coef, partials, static_cast<int>(T), static_cast<int>(B), nin);
    count_launch();
    HPC_LAUNCH_CHECK();
    return HPC_RLL_OK;
}

static inline int64_t align_up_(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// workspace: [metric (T*B f32) | partials]
size_t upgo_workspace_bytes(int64_t T, int64_t B) {
    // partials: one per scan CTA -- column tiles, or (segment, tile) pairs of the small-batch T-split
    return static_cast<size_t>(align_up_(T * B * 4, 256) + ((B + 31) / 32 + 4 * sm_count() + 80) * 8 + 256);
}

}  // namespace hpcrll

extern "C" {

int hpc_rll_upgo_forward(const float* target_output, const float* rhos, const int64_t* action,
                         const float* rewards, const float* bootstrap_values, float* loss, float* coef, int64_t T,
                         int64_t B, int64_t N, int64_t global_B, void* workspace, size_t workspace_bytes,
                         void* stream_) {
    HPC_NVTX("upgo_forward");
    using namespace hpcrll;
    cudaStream_t stream = as_stream(stream_);
    HPC_REQUIRE(T > 0 && B > 0 && N > 0, "upgo_forward: sizes must be positive (T=%lld B=%lld N=%lld)", (long long)T,
                (long long)B, (long long)N);
    HPC_REQUIRE(target_output && rhos && action && rewards && bootstrap_values && loss && coef && workspace,
                "upgo_forward: null pointer");
    HPC_REQUIRE(workspace_bytes >= upgo_workspace_bytes(T, B), "upgo_forward: workspace too small");
    HPC_REQUIRE(T < (int64_t(1) << 31) - 64 && B < (int64_t(1) << 31) - 512 && N < (int64_t(1) << 30),
                "upgo_forward: sizes exceed 2^31");
    if (global_B <= 0) global_B = B;
    const int64_t R = T * B;
    char* ws = static_cast<char*>(workspace);
    float* metric = reinterpret_cast<float*>(ws);
    double* partials = reinterpret_cast<double*>(ws + align_up_(R * 4, 256));
    const double inv_n = 1.0 / (static_cast<double>(T) * static_cast<double>(global_B));

    const RowGeom ge = row_geom(N, target_output);
    int log2G = 0;
    while ((1 << log2G) < ge.G) ++log2G;
    const bool staged = use_staged_rows(N, ge.width);
    const unsigned grid1 = rows_grid(R, staged ? kStageRows : (ge.kmax == 0 ? 8 : (32 / ge.G) * 8));
    const int n = static_cast<int>(N);
#define HPC_UP_ROWS(K, V)                                                                                       \
    do {                                                                                                        \
        if (ge.G == 1)                                                                                          \
            upgo_rows_fwd<K, V, true><<<grid1, 256, 0, stream>>>(target_output, action, metric, R, n, 1, 0);    \
        else                                                                                                    \
            upgo_rows_fwd<K, V, false><<<grid1, 256, 0, stream>>>(target_output, action, metric, R, n, ge.G,    \
                                                                  log2G);                                       \
    } while (0)
    if (staged)
        upgo_rows_fwd_staged<<<grid1, kStageRows, stage_bytes(n, 1), stream>>>(target_output, action, metric, R, n,
                                                                               stage_pitch(n),
                                                                               aligned16(target_output) ? 1 : 0);
    else if (ge.kmax == 0) upgo_rows_fwd_loop<<<grid1, 256, 0, stream>>>(target_output, action, metric, R, n);
    else HPC_ROW_DISPATCH(ge, HPC_UP_ROWS);
#undef HPC_UP_ROWS
    count_launch();
    HPC_LAUNCH_CHECK();

    const float nin = static_cast<float>(-inv_n);
    const bool tma = tma_ok_2d(bootstrap_values, B, B) && tma_ok_2d(rewards, B, B) && tma_ok_2d(rhos, B, B);
    int cfg = tuning_config(HPC_RLL_OP_UPGO);
    LbGeom lg;
    const bool lookback = lookback_geometry(HPC_RLL_OP_UPGO, T, B, &lg);
    if (!tma) cfg = 99;
    if (cfg < 0 || cfg == 21) cfg = B >= 64 * static_cast<int64_t>(sm_count()) ? 0 : 2;
    int nblocks = 0, rc = HPC_RLL_OK;
    if (lookback) {
        LbScratch sc;
        rc = lookback_scratch(lg, B, stream, &sc);
        if (rc) return rc;
        nblocks = lg.S * lg.tiles;
        upgo_scan_lookback<<<static_cast<unsigned>(nblocks), kLbCols, 0, stream>>>(
            bootstrap_values, rewards, rhos, metric, coef, partials, static_cast<int>(T), static_cast<int>(B), nin, lg.S,
            lg.L, lg.tiles, sc.ctl, sc.words);
        count_launch();
        HPC_LAUNCH_CHECK();
    } else if (cfg == 99) {
        nblocks = static_cast<int>((B + 127) / 128);
        upgo_scan_generic<<<nblocks, 128, 0, stream>>>(bootstrap_values, rewards, rhos, metric, coef, partials,
                                                       static_cast<int>(T), static_cast<int>(B), nin);
        count_launch();
        HPC_LAUNCH_CHECK();
    } else if (cfg == 0) {
        rc = launch_upgo_scan<64, 8, 4>(bootstrap_values, rewards, rhos, metric, coef, partials, T, B, nin, stream, &nblocks);
    } else if (cfg == 1) {
        rc = launch_upgo_scan<128, 8, 4>(bootstrap_values, rewards, rhos, metric, coef, partials, T, B, nin, stream, &nblocks);
    } else {
        rc = launch_upgo_scan<32, 16, 3>(bootstrap_values, rewards, rhos, metric, coef, partials, T, B, nin, stream, &nblocks);
    }
    if (rc) return rc;
    FinSpec spec;
    for (int k = 0; k < 5; ++k) spec.off[k] = spec.cnt[k] = 0, spec.scale[k] = 0.0;
    spec.cnt[0] = nblocks;
    spec.scale[0] = -inv_n;
    return launch_finalize_terms(partials, spec, 1, loss, stream);
}

int hpc_rll_upgo_backward(const float* grad_loss, const float* target_output, const int64_t* action,
                          const float* coef, float* grad_target_output, int64_t T, int64_t B, int64_t N,
                          void* stream_) {
    HPC_NVTX("upgo_backward");
    using namespace hpcrll;
    HPC_REQUIRE(T > 0 && B > 0 && N > 0, "upgo_backward: sizes must be positive");
    HPC_REQUIRE(grad_loss && target_output && action && coef && grad_target_output, "upgo_backward: null pointer");
    return launch_softmax_grad_rows(target_output, action, coef, nullptr, grad_loss, nullptr, 0.0,
                                    grad_target_output, T * B, N, false, as_stream(stream_));
}

}  // extern "C"
