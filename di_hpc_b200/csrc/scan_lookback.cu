// scan_lookback.cu -- host side of the single-launch T-split (scan_lookback.cuh): when to use it, and its scratch.
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "scan_lookback.cuh"

namespace hpcrll {

// Geometry and automatic rule, from measurements on B200 (profiles/r02_small_batch.md).  The column-scan kernels walk
// T as one dependent chain per thread: ~25 ns per step and direction whatever B is (T=1024: ~25 us per kernel for every
// B <= 4096).  A look-back launch costs ~8 us of fixed latency (launch, one load round trip, publish/fold, store), and
// its segments should fit ONE shared-memory chunk (32 rows) so that nothing is loaded twice and the two passes are 32
// steps each.  Hence: segments of 32 rows (64+ only when T > 2048, since at most kLbMaxSeg segments are folded), and
// the T-split is taken automatically when T >= 512 (below that the serial chain is already shorter than the fixed
// cost) and B <= 2048 (measured at T=1024, fwd+bwd under graph replay: B=64 21 us vs 49, B=1024 29 vs 51, B=4096 66 vs 53).  tuning config 21 forces the look-back path for
// any shape that splits, every other forced config disables it.
bool lookback_geometry(int op, int64_t T, int64_t B, LbGeom* g) {
    const int forced = tuning_config(op);
    if (forced >= 0 && forced != 21) return false;
    if (forced < 0 && (B > 2048 || T < 512)) return false;
    if (T < 32 || B <= 0) return false;
    const int64_t tiles = (B + kLbCols - 1) / kLbCols;
    int64_t L = kLbChunkRows;
    int64_t S = (T + L - 1) / L;
    if (S > kLbMaxSeg) {
        L = ((T + kLbMaxSeg - 1) / kLbMaxSeg + 7) / 8 * 8;
        S = (T + L - 1) / L;
    }
    if (S < 2) return false;
    g->S = static_cast<int>(S);
    g->L = static_cast<int>(L);
    g->tiles = static_cast<int>(tiles);
    return true;
}

namespace {
struct Entry {
    LbScratch sc;
    uint64_t last_use;
};
std::mutex g_mu;
std::map<std::pair<int, cudaStream_t>, Entry> g_scratch;
uint64_t g_tick = 0;
constexpr size_t kMaxEntries = 64;

bool capturing(cudaStream_t stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone;
}
}  // namespace

// One scratch per (device, stream): kernels that share it are ordered by the stream.  Created (cudaMalloc + blocking
// memset, epoch = 1) on first use of a stream.  That is not possible while the stream is being CAPTURED (allocation and
// the legacy-stream memset are illegal then), and PyTorch's graph helpers (`torch.cuda.graph`,
// `torch.cuda.make_graphed_callables`) capture on an internal stream the caller never sees -- so every eager call also
// keeps a few zeroed SPARE scratches per device in stock, and a capturing stream that has none adopts a spare.  The usual
// "run the step eagerly once before capturing" is therefore all a caller has to do.  Scratches are grown, never shrunk,
// and only freed (after a device synchronise, never during capture) when more than 64 streams have come by.
namespace {
constexpr size_t kSpareWords = size_t(1) << 18;  // 2 MB: covers every automatic shape (S <= 64, B <= 2048)
constexpr int kSparesPerDevice = 4;
std::map<int, std::vector<LbScratch>> g_spares;

int new_scratch(size_t cap, LbScratch* out) {
    void* p = nullptr;
    const size_t bytes = 256 + cap * sizeof(unsigned long long);
    HPC_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    LbCtl init{0u, 0u, 1u, 0u};
    if (e == cudaSuccess) e = cudaMemcpy(p, &init, sizeof(init), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(p);
        return set_error(HPC_RLL_ECUDA, "small-batch scan: scratch initialisation failed: %s", cudaGetErrorString(e));
    }
    out->ctl = static_cast<LbCtl*>(p);
    out->words = reinterpret_cast<unsigned long long*>(static_cast<char*>(p) + 256);
    out->cap_words = cap;
    return HPC_RLL_OK;
}
}  // namespace

int lookback_scratch(const LbGeom& g, int64_t B, cudaStream_t stream, LbScratch* out) {
    (void)B;
    int dev = 0;
    HPC_CUDA(cudaGetDevice(&dev));
    const size_t need = static_cast<size_t>(g.S) * static_cast<size_t>(g.tiles) * kLbCols * 2;
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_pair(dev, stream);
    auto it = g_scratch.find(key);
    if (it != g_scratch.end() && it->second.sc.cap_words >= need) {
        it->second.last_use = ++g_tick;
        *out = it->second.sc;
        return HPC_RLL_OK;
    }
    if (capturing(stream)) {
        auto& spares = g_spares[dev];
        for (size_t i = 0; it == g_scratch.end() && i < spares.size(); ++i)
            if (spares[i].cap_words >= need) {
                Entry en;
                en.sc = spares[i];
                en.last_use = ++g_tick;
                spares.erase(spares.begin() + static_cast<long>(i));
                g_scratch[key] = en;
                *out = en.sc;
                return HPC_RLL_OK;
            }
        return set_error(HPC_RLL_EINVAL,
                         "small-batch scan: no scratch for the capturing stream -- run the same call once eagerly "
                         "(any stream) before capturing a CUDA graph");
    }
    if (it != g_scratch.end()) {  // grow
        HPC_CUDA(cudaStreamSynchronize(stream));
        cudaFree(it->second.sc.ctl);
        g_scratch.erase(it);
    } else if (g_scratch.size() >= kMaxEntries) {
        auto lru = g_scratch.begin();
        for (auto i = g_scratch.begin(); i != g_scratch.end(); ++i)
            if (i->first.first == dev && (lru->first.first != dev || i->second.last_use < lru->second.last_use)) lru = i;
        if (lru->first.first == dev) {
            HPC_CUDA(cudaDeviceSynchronize());
            cudaFree(lru->second.sc.ctl);
            g_scratch.erase(lru);
        }
    }
    Entry en;
    int rc = new_scratch(need < (size_t(1) << 15) ? (size_t(1) << 15) : need, &en.sc);
    if (rc) return rc;
    en.last_use = ++g_tick;
    g_scratch[key] = en;
    *out = en.sc;
    // keep spares in stock for streams that first show up while capturing (sized for this shape if it is larger)
    auto& spares = g_spares[dev];
    while (static_cast<int>(spares.size()) < kSparesPerDevice) {
        LbScratch sp;
        if (new_scratch(need > kSpareWords ? need : kSpareWords, &sp) != HPC_RLL_OK) break;
        spares.push_back(sp);
    }
    return HPC_RLL_OK;
}

}  // namespace hpcrll
