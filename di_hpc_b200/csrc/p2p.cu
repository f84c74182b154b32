// p2p.cu -- the path's ONLY collective, done over NVLink peer memory instead of NCCL (round 2).
//
// Batch-sharded loss ops (di_hpc_b200/sharding.py) end in an all-reduce(SUM) of <= 16 scalars.  NCCL moves those 4-64 bytes
// with a ~25-50 us kernel + proxy round trip per call (profiles/r02_scaling.md: +28 us at 2 ranks, +48 us at 8 on a 227 us
// op).  Here every rank owns a small device buffer that its peers have mapped (CUDA IPC, NVLink P2P); one 32-thread
// kernel per rank
//   1. stores its scalars, each packed with the call's epoch into ONE 64-bit word (value and "ready" tag can never be
//      seen torn, no fence needed -- the same trick as scan_lookback.cuh), into slot [epoch parity][my rank] of EVERY
//      peer's buffer (system-scope stores over NVLink),
//   2. polls its OWN buffer until all `world` slots carry this epoch,
//   3. sums them in rank order -- every rank adds the same numbers in the same order, so the result is bit-identical
//      everywhere and run to run -- and bumps its epoch (device-resident: survives CUDA-graph replay).
// Two slot sets alternate by epoch parity: a rank can only be one call ahead of a peer (call k+1 cannot finish before the
// peer has entered it), so the set being overwritten was read by everybody.  The reference has no multi-GPU path at all
// (SURVEY.md 2.3).
#include "common.cuh"

namespace hpcrll {

constexpr int kP2PMaxScalars = 16;
constexpr int kP2PMaxWorld = 64;
// buffer layout: [0] epoch (u32) ... [256 B header] then 2 parity sets x kP2PMaxWorld ranks x kP2PMaxScalars u64 words
constexpr size_t kP2PBytes = 256 + sizeof(unsigned long long) * 2 * kP2PMaxWorld * kP2PMaxScalars;

struct P2PPeers {
    unsigned long long* slot[kP2PMaxWorld];  // slot area of every rank's buffer, [rank] = mine
};

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(32) allreduce_scalars_p2p_kernel(float* __restrict__ vals, int n, P2PPeers peers,
                                                                   unsigned* __restrict__ epoch_ptr, int rank, int world) {
    const int lane = threadIdx.x;
    unsigned epoch;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(epoch) : "l"(epoch_ptr) : "memory");
    const size_t set = static_cast<size_t>(epoch & 1u) * kP2PMaxWorld * kP2PMaxScalars;
    // 1. publish: lane i < n owns scalar i and writes it to every rank (its own slot included)
    float mine = 0.f;
    if (lane < n) {
        mine = vals[lane];
        const unsigned long long w = (static_cast<unsigned long long>(epoch) << 32) | __float_as_uint(mine);
        for (int r = 0; r < world; ++r) st_sys_u64(peers.slot[r] + set + static_cast<size_t>(rank) * kP2PMaxScalars + lane, w);
    }
    // 2 + 3. gather and add in rank order
    if (lane < n) {
        const unsigned long long* my = peers.slot[rank] + set + lane;
        float s = 0.f;
        for (int r = 0; r < world; ++r) {
            unsigned long long w = ld_sys_u64(my + static_cast<size_t>(r) * kP2PMaxScalars);
            while (static_cast<unsigned>(w >> 32) != epoch) w = ld_sys_u64(my + static_cast<size_t>(r) * kP2PMaxScalars);
            s = __fadd_rn(s, __uint_as_float(static_cast<unsigned>(w)));
        }
        vals[lane] = s;
    }
    __syncwarp();
    if (lane == 0) {
        unsigned e = epoch + 1u;
        if (e == 0u) e = 1u;  // 0 is the tag of a fresh (zeroed) buffer
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(epoch_ptr), "r"(e) : "memory");
    }
}

}  // namespace hpcrll

extern "C" {

size_t hpc_rll_p2p_buffer_bytes(void) { return hpcrll::kP2PBytes; }

int hpc_rll_p2p_alloc(void** local_buf, void* ipc_handle_64) {
    using namespace hpcrll;
    HPC_REQUIRE(local_buf && ipc_handle_64, "p2p_alloc: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    void* p = nullptr;
    HPC_CUDA(cudaMalloc(&p, kP2PBytes));
    cudaError_t e = cudaMemset(p, 0, kP2PBytes);
    const unsigned one = 1u;
    if (e == cudaSuccess) e = cudaMemcpy(p, &one, sizeof(one), cudaMemcpyHostToDevice);  // epoch starts at 1
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return set_error(HPC_RLL_ECUDA, "p2p_alloc: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle_64, &h, sizeof(h));
    *local_buf = p;
    return HPC_RLL_OK;
}

int hpc_rll_p2p_open(const void* ipc_handle_64, void** peer_buf) {
    using namespace hpcrll;
    HPC_REQUIRE(ipc_handle_64 && peer_buf, "p2p_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle_64, sizeof(h));
    HPC_CUDA(cudaIpcOpenMemHandle(peer_buf, h, cudaIpcMemLazyEnablePeerAccess));
    return HPC_RLL_OK;
}

int hpc_rll_p2p_close(void* peer_buf) {
    using namespace hpcrll;
    if (peer_buf) HPC_CUDA(cudaIpcCloseMemHandle(peer_buf));
    return HPC_RLL_OK;
}

int hpc_rll_p2p_free(void* local_buf) {
    using namespace hpcrll;
    if (local_buf) HPC_CUDA(cudaFree(local_buf));
    return HPC_RLL_OK;
}

int hpc_rll_allreduce_scalars_p2p(float* vals, int n, void* const* bufs, int rank, int world, void* stream) {
    using namespace hpcrll;
    HPC_NVTX("allreduce_scalars_p2p");
    HPC_REQUIRE(vals && bufs, "allreduce_scalars_p2p: null pointer");
    HPC_REQUIRE(n >= 1 && n <= kP2PMaxScalars, "allreduce_scalars_p2p: 1..%d scalars, got %d", kP2PMaxScalars, n);
    HPC_REQUIRE(world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, "allreduce_scalars_p2p: bad rank/world %d/%d",
                rank, world);
    P2PPeers peers;
    for (int r = 0; r < world; ++r) {
        HPC_REQUIRE(bufs[r] != nullptr, "allreduce_scalars_p2p: buffer of rank %d missing", r);
        peers.slot[r] = reinterpret_cast<unsigned long long*>(static_cast<char*>(bufs[r]) + 256);
    }
    for (int r = world; r < kP2PMaxWorld; ++r) peers.slot[r] = nullptr;
    allreduce_scalars_p2p_kernel<<<1, 32, 0, as_stream(stream)>>>(vals, n, peers, static_cast<unsigned*>(bufs[rank]), rank,
                                                                  world);
    count_launch();
    HPC_LAUNCH_CHECK();
    return HPC_RLL_OK;
}

}  // extern "C"
