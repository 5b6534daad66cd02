/*
 * hb_comm.cu -- the multi-GPU entry points of the C ABI (include/huffman_b200.h, "multi-GPU" section).
 *
 * The reference is single-GPU (hist.cu:67 hard-codes device 0): this part has no reference
 * equivalent.  One process (or host thread) per GPU, one hb_ctx and one hb_comm each.  A contiguous shard per rank;
 * NCCL over NVLink carries exactly two tiny collectives,
 *     1. ncclAllReduce(sum) of the 256-bin histogram      -> every rank builds the identical codebook on its host,
 *     2. ncclAllGather of one uint64 per rank: the shard's bit total sum_s hist_r[s] * len[s], known BEFORE encoding
 *                                                          -> exclusive prefix = the shard's global start bit,
 * and each rank then encodes with start_bit = offset mod 32, so its words are already in global phase
 * (hb_shard_encode_async).  The data path has no collective.
 *
 * The optional stitch gathers the shards into ONE stream on a root GPU without a funnel: the root's buffer is
 * exported with CUDA IPC, every rank maps it and PUSHES its own words straight to their final place with peer stores
 * over NVLink (stitch_push_kernel), all ranks at the same time.  A seam word (shared by neighbouring shards) is written
 * by the lowest rank that owns bits in it, OR-ed with the head words of the others (one all-gather of a word per rank).
 *
 * NCCL is bound at run time (dlopen of the libnccl.so.2 the process already has, else the system one): libhuffb200
 * itself links no NCCL and single-GPU callers never touch it.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/huffman_b200.h"
#include "hb_ctx.h"
#include "hb_kernels.cuh"

static_assert(HB_UNIQUE_ID_BYTES == sizeof(ncclUniqueId), "hb_comm_unique_id carries an ncclUniqueId");

namespace {

struct Nccl {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok = false;
};

Nccl *nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("HB_NCCL_LIB");
        if (env) n.handle = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        // the NCCL this process already uses (e.g. the one torch.distributed loaded), so that there is only one
        if (!n.handle) n.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!n.handle) n.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!n.handle) n.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!n.handle) return;
#define HB_SYM(name) n.name = reinterpret_cast<decltype(n.name)>(dlsym(n.handle, "nccl" #name))
        HB_SYM(GetUniqueId);
        HB_SYM(CommInitRank);
        HB_SYM(CommDestroy);
        HB_SYM(AllReduce);
        HB_SYM(AllGather);
        HB_SYM(Broadcast);
        HB_SYM(GetErrorString);
        HB_SYM(GetVersion);
#undef HB_SYM
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.AllGather && n.Broadcast;
    });
    return n.ok ? &n : nullptr;
}

}  // namespace

struct hb_comm {
    hb_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    bool owns_comm = false;
    int rank = 0, n_ranks = 1;
    int last_nccl = 0;

    // device scratch: [0,256) local histogram, [256,512) global histogram, [512, 512+n) gathered words
    unsigned long long *d_scratch = nullptr;
    unsigned long long *h_scratch = nullptr;     // pinned mirror
    // the plan of the last hb_shard_plan_build (the stitch needs every rank's offsets)
    uint64_t shard_bits[HB_MAX_RANKS] = {}, start_bits[HB_MAX_RANKS] = {};
    bool have_plan = false;

    // stitch target: the root's buffer, mapped here through CUDA IPC (the root uses its own pointer)
    uint32_t *stitch_base = nullptr;             // root: cudaMalloc'ed; others: cudaIpcOpenMemHandle
    uint64_t stitch_words = 0;
    int stitch_root = -1;
    bool stitch_mapped = false;
};

namespace {

#define HB_CUDA_C(c, call)                                          \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) {                                   \
            (c)->ctx->last_cuda = (int)e__;                         \
            (void)cudaGetLastError();                               \
            return HB_ERR_CUDA;                                     \
        }                                                           \
    } while (0)
#define HB_NCCL_C(c, call)                                          \
    do {                                                            \
        ncclResult_t r__ = (call);                                  \
        if (r__ != ncclSuccess) {                                   \
            (c)->last_nccl = (int)r__;                              \
            return HB_ERR_NCCL;                                     \
        }                                                           \
    } while (0)

struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) (void)cudaSetDevice(dev);
    }
    ~DeviceScope()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) (void)cudaSetDevice(prev);
    }
};

int comm_alloc(hb_comm **out, hb_ctx *ctx, int rank, int n_ranks)
{
    if (!out) return HB_ERR_ARG;
    *out = nullptr;
    if (!ctx || n_ranks < 1 || rank < 0 || rank >= n_ranks || n_ranks > HB_MAX_RANKS) return HB_ERR_ARG;
    hb_comm *c = new (std::nothrow) hb_comm();
    if (!c) return HB_ERR_NOMEM;
    c->ctx = ctx;
    c->rank = rank;
    c->n_ranks = n_ranks;
    DeviceScope g(ctx->device);
    const size_t n = 512 + 2 * (size_t)HB_MAX_RANKS;
    if (cudaMalloc(&c->d_scratch, n * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost(&c->h_scratch, n * sizeof(unsigned long long)) != cudaSuccess) {
        (void)cudaGetLastError();
        cudaFree(c->d_scratch);
        delete c;
        return HB_ERR_NOMEM;
    }
    *out = c;
    return HB_OK;
}

// ---- stitch: push this rank's words to their place in the root's stream ---------------------------------------
// dst/src are the same words in the two buffers (dst = root stream + first word of the span, src = local words +
// the same offset); both have the same 16-byte misalignment (hb_shard_local_offset), so the body goes as 128-bit
// loads and (peer) stores.  The last word of the span may be a seam word this rank owns: OR the head words of the
// ranks [or_lo, or_hi) into it.
__global__ void __launch_bounds__(512) stitch_push_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src,
                                                          unsigned long long n, const unsigned long long *heads,
                                                          int or_lo, int or_hi)
{
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    if (n == 0) return;
    // the seam word first (one thread): independent of everything else
    if (tid == 0) {
        uint32_t w = src[n - 1];
        for (int r = or_lo; r < or_hi; r++) w |= (uint32_t)heads[r];
        dst[n - 1] = w;
    }
    const unsigned long long body = n - 1;                                 // words [0, body) are plain copies
    const unsigned long long mis = ((16u - ((unsigned long long)(uintptr_t)dst & 15u)) & 15u) / 4u;
    const unsigned long long head = mis < body ? mis : body;
    if (tid < head) dst[tid] = src[tid];
    if (((uintptr_t)(src + head) & 15u) == 0) {
        const unsigned long long n_vec = (body - head) / 4;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
        unsigned long long i = tid;
        // four independent 128-bit loads in flight per thread: NVLink stores are posted, the loads are what waits
        for (; i + 3 * stride < n_vec; i += 4 * stride) {
            const uint4 a = __ldcs(s4 + i), b = __ldcs(s4 + i + stride), c = __ldcs(s4 + i + 2 * stride),
                        d = __ldcs(s4 + i + 3 * stride);
            d4[i] = a;
            d4[i + stride] = b;
            d4[i + 2 * stride] = c;
            d4[i + 3 * stride] = d;
        }
        for (; i < n_vec; i += stride) d4[i] = __ldcs(s4 + i);
        const unsigned long long done = head + n_vec * 4;
        if (tid < body - done) dst[done + tid] = src[done + tid];
    } else {
        for (unsigned long long i = head + tid; i < body; i += stride) dst[i] = src[i];   // (unmatched alignment)
    }
}

}  // namespace

extern "C" {

int hb_comm_unique_id(uint8_t id[HB_UNIQUE_ID_BYTES])
{
    if (!id) return HB_ERR_ARG;
    Nccl *n = nccl();
    if (!n) return HB_ERR_NCCL;
    ncclUniqueId u;
    if (n->GetUniqueId(&u) != ncclSuccess) return HB_ERR_NCCL;
    memcpy(id, &u, sizeof(u));
    return HB_OK;
}

int hb_comm_init(hb_comm **comm, hb_ctx *ctx, int rank, int n_ranks, const uint8_t id[HB_UNIQUE_ID_BYTES])
{
    if (!id) return HB_ERR_ARG;
    Nccl *n = nccl();
    if (!n) return HB_ERR_NCCL;
    int rc = comm_alloc(comm, ctx, rank, n_ranks);
    if (rc != HB_OK) return rc;
    DeviceScope g(ctx->device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    const ncclResult_t r = n->CommInitRank(&(*comm)->comm, n_ranks, u, rank);
    if (r != ncclSuccess) {
        hb_comm_free(*comm);
        *comm = nullptr;
        return HB_ERR_NCCL;
    }
    (*comm)->owns_comm = true;
    return HB_OK;
}

int hb_comm_adopt(hb_comm **comm, hb_ctx *ctx, void *nccl_comm, int rank, int n_ranks)
{
    if (!nccl_comm) return HB_ERR_ARG;
    if (!nccl()) return HB_ERR_NCCL;
    const int rc = comm_alloc(comm, ctx, rank, n_ranks);
    if (rc != HB_OK) return rc;
    (*comm)->comm = (ncclComm_t)nccl_comm;
    (*comm)->owns_comm = false;
    return HB_OK;
}

void hb_comm_free(hb_comm *c)
{
    if (!c) return;
    DeviceScope g(c->ctx->device);
    (void)cudaDeviceSynchronize();
    (void)hb_stitch_close(c);
    if (c->owns_comm && c->comm && nccl()) (void)nccl()->CommDestroy(c->comm);
    cudaFree(c->d_scratch);
    if (c->h_scratch) cudaFreeHost(c->h_scratch);
    (void)cudaGetLastError();
    delete c;
}

int hb_comm_last_nccl_error(const hb_comm *c) { return c ? c->last_nccl : 0; }

int hb_shard_plan_build(hb_comm *c, const uint32_t *d_in, uint64_t n_words, uint32_t codewords[256],
                        uint32_t codewordlens[256], uint64_t hist_global[256], hb_shard_plan *plan,
                        void *stream)
{
    if (!c || !codewords || !codewordlens || !plan || (!d_in && n_words)) return HB_ERR_ARG;
    Nccl *n = nccl();
    if (!n) return HB_ERR_NCCL;
    hb_ctx *ctx = c->ctx;
    DeviceScope g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *d_local = c->d_scratch, *d_global = c->d_scratch + 256, *d_gather = c->d_scratch + 512;
    unsigned long long *h_local = c->h_scratch, *h_global = c->h_scratch + 256, *h_gather = c->h_scratch + 512;

    // local histogram (hist_kernel), then collective 1: all-reduce of the 256 bins (2 KiB)
    HB_CUDA_C(c, cudaMemsetAsync(d_local, 0, 256 * sizeof(unsigned long long), st));
    int rc = hb_histogram_device(ctx, d_in, n_words, (uint64_t *)d_local, stream);
    if (rc != HB_OK) return rc;
    HB_NCCL_C(c, n->AllReduce(d_local, d_global, 256, ncclUint64, ncclSum, c->comm, st));
    HB_CUDA_C(c, cudaMemcpyAsync(h_local, d_local, 512 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    HB_CUDA_C(c, cudaStreamSynchronize(st));

    // identical codebook on every rank (hb_build_codebook is deterministic), then this shard's exact bit total
    const int max_len = hb_build_codebook((const uint64_t *)h_global, codewords, codewordlens);
    if (max_len < 0) return max_len;
    if (hist_global) memcpy(hist_global, h_global, 256 * sizeof(uint64_t));
    const uint64_t my_bits = hb_bits_from_hist((const uint64_t *)h_local, codewordlens);

    // collective 2: all-gather of one uint64 per rank
    h_gather[HB_MAX_RANKS] = my_bits;
    HB_CUDA_C(c, cudaMemcpyAsync(d_gather + HB_MAX_RANKS, h_gather + HB_MAX_RANKS, sizeof(unsigned long long),
                                 cudaMemcpyHostToDevice, st));
    HB_NCCL_C(c, n->AllGather(d_gather + HB_MAX_RANKS, d_gather, 1, ncclUint64, c->comm, st));
    HB_CUDA_C(c, cudaMemcpyAsync(h_gather, d_gather, (size_t)c->n_ranks * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, st));
    HB_CUDA_C(c, cudaStreamSynchronize(st));

    for (int r = 0; r < c->n_ranks; r++) c->shard_bits[r] = h_gather[r];
    uint64_t total = 0;
    rc = hb_shard_offsets(c->shard_bits, c->n_ranks, c->start_bits, &total);
    if (rc != HB_OK) return rc;
    c->have_plan = true;

    memset(plan, 0, sizeof(*plan));
    plan->rank = c->rank;
    plan->n_ranks = c->n_ranks;
    plan->max_len = max_len;
    plan->shard_bits = my_bits;
    plan->start_bit = c->start_bits[c->rank];
    plan->total_bits = total;
    plan->phase = (uint32_t)(plan->start_bit & 31u);
    plan->first_word = plan->start_bit >> 5;
    plan->local_words = (plan->phase + my_bits + 31) / 32;
    if (plan->local_words == 0) plan->local_words = 1;
    plan->local_offset_words = (uint32_t)(plan->first_word & 3u);
    return HB_OK;
}

int hb_shard_encode_async(hb_comm *c, const uint32_t *d_in, uint64_t n_words, const uint32_t codewords[256],
                          const uint32_t codewordlens[256], uint32_t *d_local, uint64_t local_capacity_words,
                          const hb_shard_plan *plan, void *stream)
{
    if (!c || !plan || !d_local) return HB_ERR_ARG;
    if (local_capacity_words <= plan->local_offset_words) return HB_ERR_CAPACITY;
    // the shard's words start `local_offset_words` into the buffer: same 16-byte phase as their place in the stream
    return hb_encode_async(c->ctx, d_in, n_words, codewords, codewordlens, d_local + plan->local_offset_words,
                           local_capacity_words - plan->local_offset_words, plan->phase, stream);
}

// Fused encode + stitch: the shard is encoded STRAIGHT into the root's stream (the kernel's copy-out stores go to peer
// memory over NVLink, words already in global phase): no local copy of the output, no second pass.  The two words a shard
// may share with its neighbours are OR-ed (system-scope reductions) into words the root has zeroed; the kernel leaves
// the word after a word-aligned end to the next shard.
int hb_shard_encode_direct_async(hb_comm *c, const uint32_t *d_in, uint64_t n_words, const uint32_t codewords[256],
                                 const uint32_t codewordlens[256], const hb_shard_plan *plan, void *stream)
{
    if (!c || !plan) return HB_ERR_ARG;
    if (!c->stitch_base || !c->have_plan) return HB_ERR_STATE;
    Nccl *n = nccl();
    if (!n) return HB_ERR_NCCL;
    hb_ctx *ctx = c->ctx;
    DeviceScope g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int R = c->n_ranks, me = c->rank;
    unsigned long long *d_mine = c->d_scratch + 512 + HB_MAX_RANKS;

    // the root zeroes every seam word (the first word of a shard that starts mid-word); nobody stores before that
    if (me == c->stitch_root)
        for (int r = 0; r < R; r++)
            if (c->shard_bits[r] && (c->start_bits[r] & 31u))
                HB_CUDA_C(c, cudaMemsetAsync(c->stitch_base + (c->start_bits[r] >> 5), 0, sizeof(uint32_t), st));
    HB_CUDA_C(c, cudaMemsetAsync(d_mine, 0, sizeof(unsigned long long), st));
    HB_NCCL_C(c, n->AllReduce(d_mine, d_mine, 1, ncclUint64, ncclSum, c->comm, st));

    const uint64_t bits = c->shard_bits[me], start = c->start_bits[me];
    int rc = HB_OK;
    if (bits) {
        bool is_last = true;
        for (int r = me + 1; r < R; r++) is_last = is_last && c->shard_bits[r] == 0;
        uint32_t flags = 0;
        if (start & 31u) flags |= hb::kSeamFirst;
        if (!is_last) flags |= hb::kSeamNoZeroWord | (((start + bits) & 31u) ? hb::kSeamLast : 0u);
        const uint64_t first = start >> 5;
        if (first >= c->stitch_words) {
            rc = HB_ERR_CAPACITY;              // (reported after the closing collective: the other ranks are in it)
        } else {
            ctx->next_seam_flags = flags;
            rc = hb_encode_async(ctx, d_in, n_words, codewords, codewordlens, c->stitch_base + first,
                                 c->stitch_words - first, start & 31u, stream);
        }
    } else {
        ctx->pending = true;                   // nothing to write; hb_shard_encode_result reports 0 bits
        ctx->pending_empty = true;
        ctx->pending_start_bit = start & 31u;
    }
    // completion: the root's stream learns that every rank's kernel (and its peer stores) has ended
    HB_CUDA_C(c, cudaMemsetAsync(d_mine, 0, sizeof(unsigned long long), st));
    HB_NCCL_C(c, n->AllReduce(d_mine, d_mine, 1, ncclUint64, ncclSum, c->comm, st));
    return rc;
}

int hb_shard_encode_result(hb_comm *c, uint64_t *shard_bits, void *stream)
{
    if (!c) return HB_ERR_ARG;
    return hb_encode_result(c->ctx, shard_bits, stream);
}

int hb_comm_plan_offsets(const hb_comm *c, uint64_t *start_bits, uint64_t *shard_bits)
{
    if (!c || !c->have_plan) return HB_ERR_STATE;
    for (int r = 0; r < c->n_ranks; r++) {
        if (start_bits) start_bits[r] = c->start_bits[r];
        if (shard_bits) shard_bits[r] = c->shard_bits[r];
    }
    return HB_OK;
}

// ---- stitch ------------------------------------------------------------------------------------------------------
int hb_stitch_open(hb_comm *c, uint64_t capacity_words, int root, uint32_t **d_stitched, void *stream)
{
    if (!c || root < 0 || root >= c->n_ranks || capacity_words == 0) return HB_ERR_ARG;
    Nccl *n = nccl();
    if (!n) return HB_ERR_NCCL;
    DeviceScope g(c->ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = hb_stitch_close(c);
    if (rc != HB_OK) return rc;

    // the root allocates (a plain cudaMalloc: IPC handles cover whole allocations) and broadcasts the handle
    unsigned char *d_handle = reinterpret_cast<unsigned char *>(c->d_scratch + 512);
    unsigned char *h_handle = reinterpret_cast<unsigned char *>(c->h_scratch + 512);
    static_assert(sizeof(cudaIpcMemHandle_t) <= 2 * HB_MAX_RANKS * sizeof(unsigned long long), "handle fits the scratch");
    if (c->rank == root) {
        HB_CUDA_C(c, cudaMalloc(&c->stitch_base, (capacity_words + 8) * sizeof(uint32_t)));
        cudaIpcMemHandle_t h;
        HB_CUDA_C(c, cudaIpcGetMemHandle(&h, c->stitch_base));
        memcpy(h_handle, &h, sizeof(h));
        HB_CUDA_C(c, cudaMemcpyAsync(d_handle, h_handle, sizeof(h), cudaMemcpyHostToDevice, st));
    }
    HB_NCCL_C(c, n->Broadcast(d_handle, d_handle, sizeof(cudaIpcMemHandle_t), ncclUint8, root, c->comm, st));
    if (c->rank != root) {
        HB_CUDA_C(c, cudaMemcpyAsync(h_handle, d_handle, sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost, st));
        HB_CUDA_C(c, cudaStreamSynchronize(st));
        cudaIpcMemHandle_t h;
        memcpy(&h, h_handle, sizeof(h));
        void *p = nullptr;
        HB_CUDA_C(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->stitch_base = (uint32_t *)p;
        c->stitch_mapped = true;
    } else {
        HB_CUDA_C(c, cudaStreamSynchronize(st));
    }
    c->stitch_words = capacity_words;
    c->stitch_root = root;
    if (d_stitched) *d_stitched = c->stitch_base;     // on the root: the stream; elsewhere: its peer mapping
    return HB_OK;
}

int hb_stitch_push(hb_comm *c, const uint32_t *d_local, const hb_shard_plan *plan, void *stream)
{
    if (!c || !plan || !d_local) return HB_ERR_ARG;
    if (!c->stitch_base || !c->have_plan) return HB_ERR_STATE;
    Nccl *n = nccl();
    if (!n) return HB_ERR_NCCL;
    DeviceScope g(c->ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int R = c->n_ranks, me = c->rank;
    const uint32_t *words = d_local + plan->local_offset_words;

    // every rank's head word (the first word of its local stream), one all-gather of 8 bytes per rank
    unsigned long long *d_heads = c->d_scratch + 512, *d_mine = c->d_scratch + 512 + HB_MAX_RANKS;
    HB_CUDA_C(c, cudaMemsetAsync(d_mine, 0, sizeof(unsigned long long), st));
    if (c->shard_bits[me])
        HB_CUDA_C(c, cudaMemcpyAsync(d_mine, words, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    HB_NCCL_C(c, n->AllGather(d_mine, d_heads, 1, ncclUint64, c->comm, st));

    // Which words do I write?  Word g of the stream belongs to the LOWEST rank with bits in it.  My span is
    // [first, last]; my first word belongs to an earlier rank iff I start mid-word (phase != 0: the bits before mine
    // are somebody's); my last word may be shared with later ranks that start mid-word inside it: I OR their heads in.
    const uint64_t bits = c->shard_bits[me];
    int rc = HB_OK;
    if (bits) {
        const uint64_t start = c->start_bits[me];
        const uint64_t first = start >> 5, last = (start + bits - 1) >> 5;
        const uint64_t skip = (start & 31u) ? 1 : 0;
        int or_lo = me + 1, or_hi = me + 1;
        for (int r = me + 1; r < R; r++) {
            if (c->shard_bits[r] == 0) { if (or_hi == r) or_hi = r + 1; continue; }   // (an empty shard's head is 0)
            if ((c->start_bits[r] >> 5) == last && (c->start_bits[r] & 31u)) or_hi = r + 1;
            else break;
        }
        // the stream's last rank also carries the reference's courtesy zero word after a word-aligned end
        bool is_last = true;
        for (int r = me + 1; r < R; r++) is_last = is_last && c->shard_bits[r] == 0;
        uint64_t n_words = last - first + 1;
        if (is_last && ((start + bits) & 31u) == 0) n_words++;                 // (then `last + 1` is the zero word)
        if (first + n_words > c->stitch_words) {
            rc = HB_ERR_CAPACITY;              // (reported after the closing collective: the other ranks are in it)
        } else if (n_words > skip) {
            const uint64_t cnt = n_words - skip;
            const bool seam_last = !(is_last && ((start + bits) & 31u) == 0);   // the zero word has no sharers
            unsigned long long want = (cnt / 4 + 511) / 512;
            if (want < 1) want = 1;
            static const char *env_bps = getenv("HB_STITCH_BLOCKS_PER_SM");          // tuning aid (default 4)
            const unsigned long long cap = (unsigned long long)c->ctx->sm_count * (env_bps && atoi(env_bps) > 0 ? atoi(env_bps) : 4);
            const unsigned grid = (unsigned)(want < cap ? want : cap);
            stitch_push_kernel<<<grid, 512, 0, st>>>(c->stitch_base + first + skip, words + skip, cnt, d_heads,
                                                     seam_last ? or_lo : 0, seam_last ? or_hi : 0);
            HB_CUDA_C(c, cudaGetLastError());
            c->ctx->launches++;
        }
    }
    // completion: a kernel's peer stores are visible when it has ended; the root learns that all ranks have ended
    // from a tiny all-reduce queued behind every rank's push
    HB_CUDA_C(c, cudaMemsetAsync(d_mine, 0, sizeof(unsigned long long), st));
    HB_NCCL_C(c, n->AllReduce(d_mine, d_mine, 1, ncclUint64, ncclSum, c->comm, st));
    return rc;
}

int hb_stitch_close(hb_comm *c)
{
    if (!c) return HB_ERR_ARG;
    if (!c->stitch_base) return HB_OK;
    DeviceScope g(c->ctx->device);
    (void)cudaDeviceSynchronize();
    // importers unmap first; the exporter may free only after every importer has (collective: a tiny all-reduce)
    if (c->stitch_mapped) (void)cudaIpcCloseMemHandle(c->stitch_base);
    Nccl *n = nccl();
    if (n && c->comm) {
        unsigned long long *d_mine = c->d_scratch + 512 + HB_MAX_RANKS;
        (void)cudaMemsetAsync(d_mine, 0, sizeof(unsigned long long), nullptr);
        (void)n->AllReduce(d_mine, d_mine, 1, ncclUint64, ncclSum, c->comm, nullptr);
        (void)cudaDeviceSynchronize();
    }
    if (!c->stitch_mapped) (void)cudaFree(c->stitch_base);
    (void)cudaGetLastError();
    c->stitch_base = nullptr;
    c->stitch_mapped = false;
    c->stitch_words = 0;
    c->stitch_root = -1;
    return HB_OK;
}

}  // extern "C"
