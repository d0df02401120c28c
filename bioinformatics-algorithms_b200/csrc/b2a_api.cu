// b2a_api.cu -- C ABI (include/b2align.h) over the sm_100a kernels.
//
// Host side of the batch boundary that replaces the loop hw2.cpp:328-338: validates the batch,
// discovers the pattern alphabet on the device, groups pairs of identical shape into pair-pairs
// (the two 16-bit halves of the s16x2 kernels), sizes the HBM record, launches fill + traceback
// per rows-per-lane class and returns one record per pair.  No CPU alignment code lives here:
// if CUDA is unavailable every compute entry point fails with B2A_ERR_CUDA.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/b2align.h"
#include "b2a_format.h"
#include "short16_fill.cuh"
#include "traceback.cuh"
#include "wide32.cuh"
#include "microbench.cuh"

using namespace b2a;

static_assert(sizeof(PairResult) == sizeof(b2a_result) && sizeof(b2a_result) == 32, "result record layout");
static_assert(B2A_OP_M == OP_M && B2A_OP_D == OP_D && B2A_OP_I == OP_I, "op codes");

namespace {

template <class T>
struct DevBuf {                                   // grow-only device buffer, reused across batches
    T* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct ClassRange { int R; uint32_t first, count, max_n; };

// host-side state of the wide32 family for the current batch
struct WideState {
    std::vector<WidePair> pairs;
    std::vector<WideTask> tasks;
    uint64_t chunks = 0, bound_ints = 0, rowbest_words = 0, prog_words = 0;
    int K = 32;
    bool alpha4 = false, store = true;
    DevBuf<WidePair> d_pairs;
    DevBuf<WideTask> d_tasks;
    DevBuf<Chunk> d_codes;
    DevBuf<int32_t> d_bound, d_final;
    DevBuf<uint32_t> d_rowbest, d_progress;       // d_progress[prog_words] is the ticket counter
    void release() {
        d_pairs.release(); d_tasks.release(); d_codes.release(); d_bound.release(); d_final.release();
        d_rowbest.release(); d_progress.release();
    }
};

} // namespace

struct b2a_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;

    // batch state
    bool have_batch = false, ran = false;
    b2a_params prm{};
    uint64_t n_pairs = 0;
    Short16Plan plan{0, 0, 0};
    int K = 0;
    uint8_t sym[4] = {0, 0, 0, 0};
    int nsym = 0;
    int tb_opt = 0;                               // walker tuning bits (B2A_TB_OPT overrides, for experiments)
    bool alpha4 = false;                          // pattern alphabet of the batch has <= 4 symbols
    std::vector<ClassRange> classes;
    std::vector<uint32_t> wide_pairs;             // pairs served by the wide32 family
    std::vector<uint64_t> h_pat_off, h_txt_off, h_ops_off;
    uint64_t total_ops_words = 0;
    uint64_t cells = 0, fill_bytes = 0, launches = 0, h2d = 0, d2h = 0;

    DevBuf<uint8_t> d_pat, d_txt;
    DevBuf<uint64_t> d_pat_off, d_txt_off, d_code_off, d_ops_off;
    DevBuf<PPDesc> d_pps;
    DevBuf<Chunk> d_codes;
    DevBuf<uint32_t> d_rowbest, d_ops, d_mask;
    DevBuf<PairResult> d_results;
    WideState wide;
};

namespace {

int fail(b2a_ctx* c, int code, const std::string& msg) { if (c) c->err = msg; return code; }
int cuda_fail(b2a_ctx* c, cudaError_t e, const char* where) {
    return fail(c, e == cudaErrorMemoryAllocation ? B2A_ERR_NOMEM : B2A_ERR_CUDA,
                std::string(where) + ": " + cudaGetErrorString(e));
}
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); } while (0)

// 256-bit presence mask of the bytes in [p, p+n)
__global__ void alphabet_kernel(const uint8_t* __restrict__ p, uint64_t n, uint32_t* __restrict__ mask) {
    __shared__ uint32_t s[8];
    if (threadIdx.x < 8) s[threadIdx.x] = 0;
    __syncthreads();
    uint32_t loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t b = p[i];
        loc[b >> 5] |= 1u << (b & 31);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t v = loc[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicOr(&s[k], v);
    }
    __syncthreads();
    if (threadIdx.x < 8 && s[threadIdx.x]) atomicOr(&mask[threadIdx.x], s[threadIdx.x]);
}

template <int R, int K>
cudaError_t launch_fill_rk(bool local, const FillArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)FILL_WARPS * a.tbl_cap * sizeof(uint2);
    const unsigned grid = (a.n_pp + FILL_WARPS - 1) / FILL_WARPS;
    cudaError_t e;
    if (local) {
        e = cudaFuncSetAttribute(short16_fill_kernel<R, K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        short16_fill_kernel<R, K, true><<<grid, FILL_WARPS * 32, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(short16_fill_kernel<R, K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        short16_fill_kernel<R, K, false><<<grid, FILL_WARPS * 32, smem, st>>>(a);
    }
    return cudaGetLastError();
}
template <int K>
cudaError_t launch_fill_k(int R, bool local, const FillArgs& a, cudaStream_t st) {
    switch (R) {
        case 1: return launch_fill_rk<1, K>(local, a, st);
        case 2: return launch_fill_rk<2, K>(local, a, st);
        case 3: return launch_fill_rk<3, K>(local, a, st);
        case 4: return launch_fill_rk<4, K>(local, a, st);
        case 5: return launch_fill_rk<5, K>(local, a, st);
        case 6: return launch_fill_rk<6, K>(local, a, st);
        case 7: return launch_fill_rk<7, K>(local, a, st);
        case 8: return launch_fill_rk<8, K>(local, a, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_fill(int K, int R, bool local, const FillArgs& a, cudaStream_t st) {
    switch (K) {
        case 2: return launch_fill_k<2>(R, local, a, st);
        case 4: return launch_fill_k<4>(R, local, a, st);
        case 8: return launch_fill_k<8>(R, local, a, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_tb(int K, bool local, const TbArgs& a, cudaStream_t st) {
    const unsigned threads = 128, grid = (2u * a.n_pp + threads - 1) / threads;
    switch (K * 2 + (local ? 1 : 0)) {
        case 4:  short16_traceback_kernel<2, false><<<grid, threads, 0, st>>>(a); break;
        case 5:  short16_traceback_kernel<2, true><<<grid, threads, 0, st>>>(a); break;
        case 8:  short16_traceback_kernel<4, false><<<grid, threads, 0, st>>>(a); break;
        case 9:  short16_traceback_kernel<4, true><<<grid, threads, 0, st>>>(a); break;
        case 16: short16_traceback_kernel<8, false><<<grid, threads, 0, st>>>(a); break;
        case 17: short16_traceback_kernel<8, true><<<grid, threads, 0, st>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

inline int chunk_steps(int K) { return 3 * (16 / K); }

// ---------------------------------------------------------------------------------------------
// wide32 family: host planning and launches
// ---------------------------------------------------------------------------------------------
template <int K, bool LOCAL, bool STORE>
cudaError_t launch_wide_fill_a(bool alpha4, const WideArgs& a, unsigned grid, cudaStream_t st) {
    if (alpha4) wide32_fill_kernel<K, LOCAL, STORE, true><<<grid, WIDE_WARPS * 32, 0, st>>>(a);
    else wide32_fill_kernel<K, LOCAL, STORE, false><<<grid, WIDE_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}
template <int K>
cudaError_t launch_wide_fill_k(bool local, bool store, bool alpha4, const WideArgs& a, unsigned grid, cudaStream_t st) {
    if (local) return store ? launch_wide_fill_a<K, true, true>(alpha4, a, grid, st) : launch_wide_fill_a<K, true, false>(alpha4, a, grid, st);
    return store ? launch_wide_fill_a<K, false, true>(alpha4, a, grid, st) : launch_wide_fill_a<K, false, false>(alpha4, a, grid, st);
}
cudaError_t launch_wide_fill(int K, bool local, bool store, bool alpha4, const WideArgs& a, unsigned grid, cudaStream_t st) {
    switch (K) {
        case 2:  return launch_wide_fill_k<2>(local, store, alpha4, a, grid, st);
        case 4:  return launch_wide_fill_k<4>(local, store, alpha4, a, grid, st);
        case 8:  return launch_wide_fill_k<8>(local, store, alpha4, a, grid, st);
        case 16: return launch_wide_fill_k<16>(local, store, alpha4, a, grid, st);
        case 32: return launch_wide_fill_k<32>(local, store, alpha4, a, grid, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_wide_tb(int K, bool local, const WideTbArgs& a, cudaStream_t st) {
    const unsigned threads = 64, grid = (a.n_wide + threads - 1) / threads;
    switch (K * 2 + (local ? 1 : 0)) {
        case 4:  wide32_traceback_kernel<2, false><<<grid, threads, 0, st>>>(a); break;
        case 5:  wide32_traceback_kernel<2, true><<<grid, threads, 0, st>>>(a); break;
        case 8:  wide32_traceback_kernel<4, false><<<grid, threads, 0, st>>>(a); break;
        case 9:  wide32_traceback_kernel<4, true><<<grid, threads, 0, st>>>(a); break;
        case 16: wide32_traceback_kernel<8, false><<<grid, threads, 0, st>>>(a); break;
        case 17: wide32_traceback_kernel<8, true><<<grid, threads, 0, st>>>(a); break;
        case 32: wide32_traceback_kernel<16, false><<<grid, threads, 0, st>>>(a); break;
        case 33: wide32_traceback_kernel<16, true><<<grid, threads, 0, st>>>(a); break;
        case 64: wide32_traceback_kernel<32, false><<<grid, threads, 0, st>>>(a); break;
        case 65: wide32_traceback_kernel<32, true><<<grid, threads, 0, st>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int wide_plan(b2a_ctx* ctx, const uint64_t* pat_off, const uint64_t* txt_off, bool store)
{
    WideState& W = ctx->wide;
    const b2a_params& prm = ctx->prm;
    W.pairs.clear(); W.tasks.clear();
    W.chunks = W.bound_ints = W.rowbest_words = W.prog_words = 0;
    W.K = delta_bits_wide(prm.match, prm.mismatch, prm.gap);
    W.store = store;
    W.alpha4 = ctx->alpha4 && prm.match <= 127 && prm.match >= -128 && prm.mismatch <= 127 && prm.mismatch >= -128;
    const int CS = 2 * (32 / W.K);
    uint32_t max_bands = 0;
    for (uint32_t k : ctx->wide_pairs) {
        WidePair p{};
        p.pat_off = pat_off[k]; p.txt_off = txt_off[k];
        p.m = (uint32_t)(pat_off[k + 1] - pat_off[k]); p.n = (uint32_t)(txt_off[k + 1] - txt_off[k]);
        p.pair = k;
        p.nbands = (p.m && p.n) ? (p.m + 32u * WIDE_R - 1u) / (32u * WIDE_R) : 0u;
        p.code_off = W.chunks;
        if (store) W.chunks += (uint64_t)p.nbands * WIDE_R * num_chunks(p.n, CS) * 32u;
        p.bound_stride = ((p.n + 64u) + 31u) & ~31u;
        p.bound_off = W.bound_ints; W.bound_ints += 2ull * p.bound_stride;
        p.rowbest_off = W.rowbest_words; W.rowbest_words += (uint64_t)p.nbands * 32u * WIDE_R;
        p.prog_off = W.prog_words; W.prog_words += p.nbands;
        max_bands = std::max(max_bands, p.nbands);
        W.pairs.push_back(p);
    }
    // ticket order: band-major, so the band a warp depends on always holds an earlier ticket
    std::vector<uint32_t> order(W.pairs.size());
    for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return W.pairs[a].nbands > W.pairs[b].nbands; });
    for (uint32_t b = 0; b < max_bands; ++b)
        for (uint32_t i : order) { if (W.pairs[i].nbands <= b) break; W.tasks.push_back(WideTask{i, b}); }
    cudaStream_t st = ctx->stream;
    CU(W.d_pairs.reserve(W.pairs.size())); CU(W.d_tasks.reserve(W.tasks.size()));
    CU(W.d_codes.reserve(W.chunks)); CU(W.d_bound.reserve(W.bound_ints)); CU(W.d_final.reserve(W.pairs.size()));
    CU(W.d_rowbest.reserve(W.rowbest_words)); CU(W.d_progress.reserve(W.prog_words + 1));
    CU(cudaMemcpyAsync(W.d_pairs.p, W.pairs.data(), W.pairs.size() * sizeof(WidePair), cudaMemcpyHostToDevice, st));
    if (!W.tasks.empty()) CU(cudaMemcpyAsync(W.d_tasks.p, W.tasks.data(), W.tasks.size() * sizeof(WideTask), cudaMemcpyHostToDevice, st));
    ctx->h2d += W.pairs.size() * sizeof(WidePair) + W.tasks.size() * sizeof(WideTask);
    ctx->fill_bytes += W.chunks * sizeof(Chunk);
    return B2A_OK;
}

int wide_fill(b2a_ctx* ctx, cudaStream_t st, uint64_t* launches)
{
    WideState& W = ctx->wide;
    if (W.tasks.empty()) return B2A_OK;
    const b2a_params& prm = ctx->prm;
    CU(cudaMemsetAsync(W.d_progress.p, 0, (W.prog_words + 1) * 4, st));
    WideArgs a{};
    a.pat = ctx->d_pat.p; a.txt = ctx->d_txt.p; a.pairs = W.d_pairs.p; a.tasks = W.d_tasks.p;
    a.n_tasks = (uint32_t)W.tasks.size(); a.ticket = W.d_progress.p + W.prog_words;
    a.codes = W.d_codes.p; a.bound = W.d_bound.p; a.rowbest = W.d_rowbest.p; a.progress = W.d_progress.p;
    a.final_score = W.d_final.p;
    a.match = prm.match; a.mismatch = prm.mismatch; a.gap = prm.gap;
    a.radix = W.K < 32 ? (1u << W.K) : 0u;
    for (int s = 0; s < 4; ++s) a.sym[s] = ctx->sym[s];
    a.nsym = ctx->nsym;
    const unsigned need = (unsigned)((W.tasks.size() + WIDE_WARPS - 1) / WIDE_WARPS);
    const unsigned grid = std::min<unsigned>(need, (unsigned)ctx->sm_count * 8u);
    CU(launch_wide_fill(W.K, prm.mode == B2A_MODE_LOCAL, W.store, W.alpha4, a, grid, st));
    ++*launches;
    return B2A_OK;
}

int wide_traceback(b2a_ctx* ctx, cudaStream_t st, uint64_t* launches, bool score_only)
{
    WideState& W = ctx->wide;
    if (W.pairs.empty()) return B2A_OK;
    const b2a_params& prm = ctx->prm;
    WideTbArgs a{};
    a.pat = ctx->d_pat.p; a.txt = ctx->d_txt.p; a.pairs = W.d_pairs.p; a.n_wide = (uint32_t)W.pairs.size();
    a.codes = W.d_codes.p; a.rowbest = W.d_rowbest.p; a.final_score = W.d_final.p;
    a.results = ctx->d_results.p;
    a.ops = (prm.flags & B2A_WANT_OPS) && !score_only ? ctx->d_ops.p : nullptr;
    a.ops_off = ctx->d_ops_off.p;
    a.match = prm.match; a.mismatch = prm.mismatch; a.gap = prm.gap; a.score_only = score_only ? 1 : 0;
    a.opt = ctx->tb_opt;
    CU(launch_wide_tb(W.K, prm.mode == B2A_MODE_LOCAL, a, st));
    ++*launches;
    return B2A_OK;
}

} // namespace

extern "C" {

int b2a_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

b2a_ctx* b2a_create(int device) {
    int n = b2a_device_count();
    if (device < 0 || device >= n) return nullptr;
    b2a_ctx* ctx = new (std::nothrow) b2a_ctx();
    if (!ctx) return nullptr;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        prop.major < 10 ||                                   // kernels are sm_100a only: no fallback
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx; cudaGetLastError(); return nullptr;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* e = std::getenv("B2A_TB_OPT")) ctx->tb_opt = std::atoi(e);
    for (auto& e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) { b2a_destroy(ctx); return nullptr; }
    return ctx;
}

void b2a_destroy(b2a_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->d_pat.release(); ctx->d_txt.release(); ctx->d_pat_off.release(); ctx->d_txt_off.release();
    ctx->d_code_off.release(); ctx->d_ops_off.release(); ctx->d_pps.release(); ctx->d_codes.release();
    ctx->d_rowbest.release(); ctx->d_ops.release(); ctx->d_mask.release(); ctx->d_results.release();
    ctx->wide.release();
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* b2a_last_error(const b2a_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (CUDA device unavailable?)"; }

void* b2a_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void b2a_host_free(void* p) { if (p) cudaFreeHost(p); }

int b2a_batch_upload(b2a_ctx* ctx, const b2a_params* prm, const uint8_t* pat, const uint64_t* pat_off,
                     const uint8_t* txt, const uint64_t* txt_off, uint64_t n_pairs)
{
    if (!ctx) return B2A_ERR_ARG;
    if (!prm || !pat_off || !txt_off || (prm->mode != B2A_MODE_GLOBAL && prm->mode != B2A_MODE_LOCAL))
        return fail(ctx, B2A_ERR_ARG, "b2a_batch_upload: null argument or bad mode");
    if (n_pairs > 0x7FFFFFF0ull) return fail(ctx, B2A_ERR_ARG, "b2a_batch_upload: too many pairs");
    CU(cudaSetDevice(ctx->device));
    ctx->have_batch = false; ctx->ran = false;
    ctx->prm = *prm; ctx->n_pairs = n_pairs;
    ctx->classes.clear(); ctx->wide_pairs.clear();
    ctx->cells = ctx->fill_bytes = ctx->launches = ctx->h2d = ctx->d2h = 0;
    const uint64_t pat_bytes = n_pairs ? pat_off[n_pairs] : 0, txt_bytes = n_pairs ? txt_off[n_pairs] : 0;
    if ((pat_bytes && !pat) || (txt_bytes && !txt)) return fail(ctx, B2A_ERR_ARG, "b2a_batch_upload: null sequence buffer");
    // Appendix A.8: (m+n)*max|score| must stay inside int32 (beyond that the reference itself is undefined)
    const int64_t smag = std::max<int64_t>({std::llabs((long long)prm->match), std::llabs((long long)prm->mismatch),
                                            std::llabs((long long)prm->gap)});

    cudaStream_t st = ctx->stream;
    CU(ctx->d_pat.reserve(pat_bytes + 16)); CU(ctx->d_txt.reserve(txt_bytes + 16));
    CU(ctx->d_pat_off.reserve(n_pairs + 1)); CU(ctx->d_txt_off.reserve(n_pairs + 1));
    CU(ctx->d_mask.reserve(8));
    if (pat_bytes) CU(cudaMemcpyAsync(ctx->d_pat.p, pat, pat_bytes, cudaMemcpyHostToDevice, st));
    if (txt_bytes) CU(cudaMemcpyAsync(ctx->d_txt.p, txt, txt_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_pat_off.p, pat_off, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_txt_off.p, txt_off, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    ctx->h2d += pat_bytes + txt_bytes + 2 * (n_pairs + 1) * 8;
    // pattern alphabet on the device (the PRMT score tables hold 4 symbols)
    CU(cudaMemsetAsync(ctx->d_mask.p, 0, 32, st));
    if (pat_bytes) {
        alphabet_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(ctx->d_pat.p, pat_bytes, ctx->d_mask.p);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    uint32_t mask[8];
    CU(cudaMemcpyAsync(mask, ctx->d_mask.p, 32, cudaMemcpyDeviceToHost, st));

    // ---- host planning (overlaps the copies above) ----
    ctx->h_pat_off.assign(pat_off, pat_off + n_pairs + 1);
    ctx->h_txt_off.assign(txt_off, txt_off + n_pairs + 1);
    ctx->h_ops_off.resize(n_pairs + 1);
    std::vector<PPDesc> pps[SHORT16_MAX_R + 1];
    std::unordered_map<uint64_t, uint32_t> pending;
    bool have_last = false; uint64_t last_key = 0; uint32_t last_idx = 0;
    uint64_t plan_key = ~0ull; bool plan_ok = false; Short16Plan pl{0, 0, 0};
    uint64_t opsw = 0, cells = 0;
    const bool want_ops = (prm->flags & B2A_WANT_OPS) != 0;
    for (uint64_t k = 0; k < n_pairs; ++k) {
        if (pat_off[k + 1] < pat_off[k] || txt_off[k + 1] < txt_off[k])
            return fail(ctx, B2A_ERR_ARG, "b2a_batch_upload: offsets must be non-decreasing");
        const uint64_t m64 = pat_off[k + 1] - pat_off[k], n64 = txt_off[k + 1] - txt_off[k];
        if ((m64 + n64 + 2) * (uint64_t)smag >= 0x7FFFFFFFull || m64 + n64 >= 0x7FFFFFF0ull)
            return fail(ctx, B2A_ERR_RANGE, "b2a_batch_upload: (m+n)*max|score| exceeds int32 (SURVEY.md A.8)");
        const uint32_t m = (uint32_t)m64, n = (uint32_t)n64;
        ctx->h_ops_off[k] = opsw;
        opsw += (m64 + n64 + 15) / 16 + 1;
        cells += m64 * n64;
        const uint64_t key = (m64 << 32) | n64;
        if (key != plan_key) { plan_key = key; plan_ok = short16_plan(prm->mode, m, n, prm->match, prm->mismatch, prm->gap, pl); }
        if (!plan_ok || (prm->flags & B2A_SCORE_ONLY)) { ctx->wide_pairs.push_back((uint32_t)k); continue; }
        auto emit = [&](uint32_t a, uint32_t b) { pps[pl.R].push_back(PPDesc{a, b, m, n}); };
        if (have_last && key == last_key) { emit(last_idx, (uint32_t)k); have_last = false; continue; }
        if (have_last) {
            auto it = pending.find(last_key);
            if (it != pending.end()) {
                const uint32_t lm = (uint32_t)(last_key >> 32), ln = (uint32_t)last_key;
                pps[(lm + 31) / 32].push_back(PPDesc{it->second, last_idx, lm, ln});
                pending.erase(it);
            } else pending[last_key] = last_idx;
            have_last = false;
        }
        auto it = pending.find(key);
        if (it != pending.end()) { emit(it->second, (uint32_t)k); pending.erase(it); }
        else { have_last = true; last_key = key; last_idx = (uint32_t)k; }
    }
    if (have_last) {
        auto it = pending.find(last_key);
        const uint32_t lm = (uint32_t)(last_key >> 32), ln = (uint32_t)last_key;
        if (it != pending.end()) { pps[(lm + 31) / 32].push_back(PPDesc{it->second, last_idx, lm, ln}); pending.erase(it); }
        else pending[last_key] = last_idx;
    }
    for (auto& kv : pending) {                                  // unpaired leftovers: both halves carry the same pair
        const uint32_t lm = (uint32_t)(kv.first >> 32), ln = (uint32_t)kv.first;
        pps[(lm + 31) / 32].push_back(PPDesc{kv.second, kv.second, lm, ln});
    }
    ctx->h_ops_off[n_pairs] = opsw;
    ctx->total_ops_words = opsw;
    ctx->cells = cells;
    ctx->K = delta_bits(prm->match, prm->mismatch, prm->gap);
    const int CS = ctx->K ? chunk_steps(ctx->K) : 24;

    std::vector<PPDesc> all;
    std::vector<uint64_t> code_off;
    uint64_t chunks = 0;
    for (int R = 1; R <= SHORT16_MAX_R; ++R) {
        if (pps[R].empty()) continue;
        ClassRange cr{R, (uint32_t)all.size(), (uint32_t)pps[R].size(), 0};
        for (const PPDesc& d : pps[R]) {
            all.push_back(d);
            code_off.push_back(chunks);
            chunks += (uint64_t)R * num_chunks(d.n, CS) * 32u;
            cr.max_n = std::max(cr.max_n, d.n);
        }
        ctx->classes.push_back(cr);
    }
    CU(cudaStreamSynchronize(st));                               // alphabet mask is on the host now
    ctx->nsym = 0;
    bool short_ok = true;
    for (int b = 0; b < 256 && short_ok; ++b)
        if (mask[b >> 5] & (1u << (b & 31))) { if (ctx->nsym == 4) short_ok = false; else ctx->sym[ctx->nsym++] = (uint8_t)b; }
    ctx->alpha4 = short_ok;
    if (!short_ok && !all.empty()) {
        // more than 4 distinct pattern symbols: the PRMT tables cannot hold them -> wide family for everything
        for (const PPDesc& d : all) { ctx->wide_pairs.push_back(d.a); if (d.b != d.a) ctx->wide_pairs.push_back(d.b); }
        std::sort(ctx->wide_pairs.begin(), ctx->wide_pairs.end());
        all.clear(); code_off.clear(); ctx->classes.clear(); chunks = 0;
    }
    const size_t n_pp = all.size();
    CU(ctx->d_pps.reserve(n_pp)); CU(ctx->d_code_off.reserve(n_pp));
    CU(ctx->d_codes.reserve(chunks));
    CU(ctx->d_results.reserve(n_pairs));
    if (prm->mode == B2A_MODE_LOCAL) {
        size_t rb = 0;
        for (const ClassRange& c : ctx->classes) rb = std::max(rb, (size_t)(c.first + c.count) * c.R * 32);
        CU(ctx->d_rowbest.reserve(rb));
    }
    if (want_ops) { CU(ctx->d_ops.reserve(opsw)); CU(ctx->d_ops_off.reserve(n_pairs + 1)); }
    if (n_pp) {
        CU(cudaMemcpyAsync(ctx->d_pps.p, all.data(), n_pp * sizeof(PPDesc), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->d_code_off.p, code_off.data(), n_pp * 8, cudaMemcpyHostToDevice, st));
        ctx->h2d += n_pp * (sizeof(PPDesc) + 8);
    }
    if (want_ops) {
        CU(cudaMemcpyAsync(ctx->d_ops_off.p, ctx->h_ops_off.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
        ctx->h2d += (n_pairs + 1) * 8;
    }
    ctx->fill_bytes = chunks * sizeof(Chunk);
    if (!ctx->wide_pairs.empty()) {
        int rc = wide_plan(ctx, pat_off, txt_off, !(prm->flags & B2A_SCORE_ONLY));
        if (rc != B2A_OK) return rc;
    } else { ctx->wide.pairs.clear(); ctx->wide.tasks.clear(); }
    CU(cudaStreamSynchronize(st));
    ctx->have_batch = true;
    return B2A_OK;
}

int b2a_batch_run(b2a_ctx* ctx, float* fill_ms, float* traceback_ms)
{
    if (!ctx) return B2A_ERR_ARG;
    if (!ctx->have_batch) return fail(ctx, B2A_ERR_STATE, "b2a_batch_run: no batch uploaded");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const b2a_params& prm = ctx->prm;
    const bool local = prm.mode == B2A_MODE_LOCAL, want_ops = (prm.flags & B2A_WANT_OPS) != 0;
    uint64_t launches = 0;
    CU(cudaEventRecord(ctx->ev[0], st));
    for (const ClassRange& c : ctx->classes) {
        Short16Plan pl{0, 0, 0};
        short16_plan(prm.mode, (uint32_t)c.R * 32u, c.max_n, prm.match, prm.mismatch, prm.gap, pl);   // bias for the class' largest shape
        FillArgs a{};
        a.pat = ctx->d_pat.p; a.txt = ctx->d_txt.p; a.pat_off = ctx->d_pat_off.p; a.txt_off = ctx->d_txt_off.p;
        a.pps = ctx->d_pps.p + c.first; a.code_off = ctx->d_code_off.p + c.first; a.codes = ctx->d_codes.p;
        a.rowbest = local ? ctx->d_rowbest.p + (size_t)c.first * c.R * 32 : nullptr;
        a.n_pp = c.count; a.tbl_cap = (c.max_n + 3u) & ~3u;
        a.match = prm.match; a.mismatch = prm.mismatch; a.gap = prm.gap; a.bias = pl.bias;
        a.radix = 1u << ctx->K;
        for (int s = 0; s < 4; ++s) a.sym[s] = ctx->sym[s];
        a.nsym = ctx->nsym;
        CU(launch_fill(ctx->K, c.R, local, a, st));
        ++launches;
    }
    if (!ctx->wide_pairs.empty()) { int rc = wide_fill(ctx, st, &launches); if (rc != B2A_OK) return rc; }
    CU(cudaEventRecord(ctx->ev[1], st));
    for (const ClassRange& c : ctx->classes) {
        Short16Plan pl{0, 0, 0};
        short16_plan(prm.mode, (uint32_t)c.R * 32u, c.max_n, prm.match, prm.mismatch, prm.gap, pl);
        TbArgs a{};
        a.pat = ctx->d_pat.p; a.txt = ctx->d_txt.p; a.pat_off = ctx->d_pat_off.p; a.txt_off = ctx->d_txt_off.p;
        a.pps = ctx->d_pps.p + c.first; a.code_off = ctx->d_code_off.p + c.first; a.codes = ctx->d_codes.p;
        a.rowbest = local ? ctx->d_rowbest.p + (size_t)c.first * c.R * 32 : nullptr;
        a.results = ctx->d_results.p; a.ops = want_ops ? ctx->d_ops.p : nullptr; a.ops_off = ctx->d_ops_off.p;
        a.n_pp = c.count; a.R = c.R;
        a.match = prm.match; a.mismatch = prm.mismatch; a.gap = prm.gap; a.bias = pl.bias;
        a.opt = ctx->tb_opt;
        CU(launch_tb(ctx->K, local, a, st));
        ++launches;
    }
    if (!ctx->wide_pairs.empty()) { int rc = wide_traceback(ctx, st, &launches, (prm.flags & B2A_SCORE_ONLY) != 0); if (rc != B2A_OK) return rc; }
    CU(cudaEventRecord(ctx->ev[2], st));
    CU(cudaStreamSynchronize(st));
    float f = 0, t = 0;
    CU(cudaEventElapsedTime(&f, ctx->ev[0], ctx->ev[1]));
    CU(cudaEventElapsedTime(&t, ctx->ev[1], ctx->ev[2]));
    if (fill_ms) *fill_ms = f;
    if (traceback_ms) *traceback_ms = t;
    ctx->launches += launches;
    ctx->ran = true;
    return B2A_OK;
}

int b2a_batch_download(b2a_ctx* ctx, b2a_result* results)
{
    if (!ctx) return B2A_ERR_ARG;
    if (!ctx->ran) return fail(ctx, B2A_ERR_STATE, "b2a_batch_download: batch not run");
    if (!results && ctx->n_pairs) return fail(ctx, B2A_ERR_ARG, "b2a_batch_download: null results");
    CU(cudaSetDevice(ctx->device));
    if (ctx->n_pairs) {
        CU(cudaMemcpyAsync(results, ctx->d_results.p, ctx->n_pairs * sizeof(b2a_result), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->d2h += ctx->n_pairs * sizeof(b2a_result);
    }
    return B2A_OK;
}

int b2a_align_batch(b2a_ctx* ctx, const b2a_params* prm, const uint8_t* pat, const uint64_t* pat_off,
                    const uint8_t* txt, const uint64_t* txt_off, uint64_t n_pairs, b2a_result* results)
{
    int rc = b2a_batch_upload(ctx, prm, pat, pat_off, txt, txt_off, n_pairs);
    if (rc != B2A_OK) return rc;
    rc = b2a_batch_run(ctx, nullptr, nullptr);
    if (rc != B2A_OK) return rc;
    return b2a_batch_download(ctx, results);
}

int b2a_batch_stats(const b2a_ctx* ctx, uint64_t* kernel_launches, uint64_t* cells, uint64_t* fill_bytes,
                    uint64_t* h2d_bytes, uint64_t* d2h_bytes)
{
    if (!ctx) return B2A_ERR_ARG;
    if (kernel_launches) *kernel_launches = ctx->launches;
    if (cells) *cells = ctx->cells;
    if (fill_bytes) *fill_bytes = ctx->fill_bytes;
    if (h2d_bytes) *h2d_bytes = ctx->h2d;
    if (d2h_bytes) *d2h_bytes = ctx->d2h;
    return B2A_OK;
}

int64_t b2a_fetch_ops(b2a_ctx* ctx, uint64_t pair, char* ops, uint64_t ops_cap)
{
    if (!ctx) return B2A_ERR_ARG;
    if (!ctx->ran || !(ctx->prm.flags & B2A_WANT_OPS)) return fail(ctx, B2A_ERR_STATE, "b2a_fetch_ops: run a batch with B2A_WANT_OPS first");
    if (pair >= ctx->n_pairs) return fail(ctx, B2A_ERR_ARG, "b2a_fetch_ops: pair index out of range");
    CU(cudaSetDevice(ctx->device));
    PairResult r;
    CU(cudaMemcpy(&r, ctx->d_results.p + pair, sizeof(r), cudaMemcpyDeviceToHost));
    if (r.n_ops > ops_cap) return fail(ctx, B2A_ERR_ARG, "b2a_fetch_ops: buffer too small");
    const uint64_t nw = ((uint64_t)r.n_ops + 15) / 16;
    std::vector<uint32_t> w(nw);
    if (nw) CU(cudaMemcpy(w.data(), ctx->d_ops.p + ctx->h_ops_off[pair], nw * 4, cudaMemcpyDeviceToHost));
    ctx->d2h += sizeof(r) + nw * 4;
    static const char L[4] = {'M', 'D', 'I', '?'};
    for (uint32_t t = 0; t < r.n_ops; ++t) ops[t] = L[(w[t >> 4] >> (2 * (t & 15))) & 3u];
    return (int64_t)r.n_ops;
}

int64_t b2a_copy_ops(b2a_ctx* ctx, uint32_t* ops_words, uint64_t cap_words, uint64_t* ops_off)
{
    if (!ctx) return B2A_ERR_ARG;
    if (!ctx->ran || !(ctx->prm.flags & B2A_WANT_OPS)) return fail(ctx, B2A_ERR_STATE, "b2a_copy_ops: run a batch with B2A_WANT_OPS first");
    if (ops_off) std::memcpy(ops_off, ctx->h_ops_off.data(), (ctx->n_pairs + 1) * 8);
    if (!ops_words) return (int64_t)ctx->total_ops_words;
    if (cap_words < ctx->total_ops_words) return fail(ctx, B2A_ERR_ARG, "b2a_copy_ops: buffer too small");
    CU(cudaSetDevice(ctx->device));
    if (ctx->total_ops_words) {
        CU(cudaMemcpyAsync(ops_words, ctx->d_ops.p, ctx->total_ops_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->d2h += ctx->total_ops_words * 4;
    }
    return (int64_t)ctx->total_ops_words;
}

int64_t b2a_debug_copy_record(b2a_ctx* ctx, void* chunks, uint64_t chunk_cap, void* rowbest, uint64_t rowbest_cap)
{
    if (!ctx) return B2A_ERR_ARG;
    if (!ctx->ran || ctx->classes.empty()) return fail(ctx, B2A_ERR_STATE, "b2a_debug_copy_record: no short16 record");
    CU(cudaSetDevice(ctx->device));
    const uint64_t bytes = ctx->fill_bytes;
    if (chunks) CU(cudaMemcpy(chunks, ctx->d_codes.p, std::min<uint64_t>(bytes, chunk_cap), cudaMemcpyDeviceToHost));
    if (rowbest && ctx->prm.mode == B2A_MODE_LOCAL)
        CU(cudaMemcpy(rowbest, ctx->d_rowbest.p, std::min<uint64_t>(ctx->d_rowbest.cap * 4, rowbest_cap), cudaMemcpyDeviceToHost));
    return (int64_t)bytes;
}

int b2a_microbench_int16x2(b2a_ctx* ctx, int kind, double* gops, float* sm_mhz)
{
    if (!ctx || !gops) return B2A_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    double g = 0; float mhz = 0;
    cudaError_t e = run_microbench(kind, ctx->sm_count, ctx->stream, &g, &mhz);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "microbench");
    *gops = g;
    if (sm_mhz) *sm_mhz = mhz;
    return B2A_OK;
}

} // extern "C"
