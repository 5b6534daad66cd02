/*
 * hb_encode.cu -- single-pass Huffman variable-length encode for sm_100a (B200).
 *
 * Replaces, in ONE kernel, the reference's three passes
 *     vlc_encode_kernel_sm64huff (vlc_kernel_sm64huff.cu:37-160)   per-block encode
 *     prescanArray               (scan.cu:114-231, scanLargeArray_kernel.cu:216-258)   block offsets
 *     cudaMemset + pack2         (main_test_cu.cu:162, pack_kernels.cu:19-52)   compaction
 * and produces the exact word stream of cpu_vlc_encode (cpuencode.cpp:12-46).
 *
 * Design (not a port: none of the reference's structure survives):
 *   - persistent CTAs pull 8 KiB tiles (256 threads x 32 symbols) from an atomic ticket counter;
 *   - the codebook lives in shared memory, replicated once per lane (256 x 32 words) so that the
 *     one table lookup per symbol is bank-conflict free for any symbol distribution; an entry is
 *     (cw << (32-len)) | len, which lets ONE funnel shift both make room in the 64-bit bit
 *     accumulator and merge the codeword (shf.l.wrap uses the low 5 bits of the same register);
 *   - pass 1 looks the 32 symbols of a thread up (kept in registers) and sums their lengths with
 *     dp4a; a shuffle scan gives every thread its bit offset inside the tile;
 *   - pass 2 re-walks the registers, appending into a 64-bit accumulator that is already aligned
 *     to the tile's 32-bit word grid, and stores a word to the shared staging buffer whenever a
 *     word boundary is crossed (checked every G symbols, G * max_len <= 32).  Because every thread
 *     of a full tile emits >= 32 bits, a staging word is shared by at most two neighbouring
 *     threads: the left one hands its partial tail word to the right one through a shuffle
 *     (through shared memory across warps), so no shared-memory atomics and no pre-zeroing;
 *     tiles that break the >= 32 bits rule (ragged last tile, zero-length codes) take an
 *     atomicOr path with identical results;
 *   - tile bit offsets come from a decoupled look-back over 64-bit descriptors
 *     {epoch, status, 48-bit bit count}; the aggregate is published BEFORE pass 2 so successors
 *     rarely wait, and nothing has to be reset between calls (epoch tag, monotonic tickets);
 *   - the staging buffer is copied out coalesced with one funnel shift per word to the global
 *     phase (P mod 32).  The global word that straddles two tiles is owned by the RIGHT tile,
 *     which re-derives the few (< 32) bits it needs from the symbols just before the tile instead
 *     of waiting for its neighbour: no inter-CTA data dependency, no atomics on the output, no
 *     memset of the output.
 */
#include "hb_kernels.cuh"

namespace hb {
namespace {

constexpr int kWarps = kEncThreads / 32;
constexpr int kPackedTabWords = 256 * 32;                     // lane-replicated packed entries
constexpr int kWideTabWords = 256 * 2;                        // uint2 {cw << (32-len), len}
constexpr int kPackedMaxLen = 24;
constexpr int kPackedStageWords = kTileBytes * kPackedMaxLen / 32 + 8;
constexpr int kWideStageWords = kTileBytes * 31 / 32 + 8;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// one 256-bit load per lane: a warp reads 1 KiB contiguous, streamed past L1 (LDG.E.256 on sm_100a)
__device__ __forceinline__ void ld_stream_v8(const uint32_t *p, uint32_t (&w)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]),
                   "=r"(w[7])
                 : "l"(p));
}
__device__ __forceinline__ unsigned long long pack_desc(uint32_t epoch, unsigned long long status,
                                                        unsigned long long bits)
{
    return ((unsigned long long)epoch << 50) | (status << kDescValueBits) | (bits & kDescValueMask);
}

// ---- codebook access ---------------------------------------------------------------------------
// An entry is handed around as two registers (c = left-aligned codeword bits, s = shift source whose
// low 5 bits are the length).  For the packed table both are the same register.
template <bool WIDE>
struct Tab;

template <>
struct Tab<false> {
    static constexpr int kWords = kPackedTabWords;
    static constexpr int kStageWords = kPackedStageWords;
    __device__ static __forceinline__ void fill(uint32_t *tab, const uint32_t *g, uint32_t tid)
    {
        for (uint32_t i = tid; i < kPackedTabWords; i += kEncThreads)
            tab[i] = __ldg(g + (i >> 5));
    }
    __device__ static __forceinline__ void look(const uint32_t *tab, uint32_t sym, uint32_t lane,
                                                uint32_t &c, uint32_t &s)
    {
        c = s = tab[(sym << 5) + lane];
    }
    // low byte of the entry is the length (bits 5..7 are zero because len <= 24)
    __device__ static __forceinline__ uint32_t add_len(uint32_t acc, uint32_t s)
    {
        return __dp4a(s, 1u, acc);
    }
};

template <>
struct Tab<true> {
    static constexpr int kWords = kWideTabWords;
    static constexpr int kStageWords = kWideStageWords;
    __device__ static __forceinline__ void fill(uint32_t *tab, const uint32_t *g, uint32_t tid)
    {
        for (uint32_t i = tid; i < kWideTabWords; i += kEncThreads)
            tab[i] = __ldg(g + i);
    }
    __device__ static __forceinline__ void look(const uint32_t *tab, uint32_t sym, uint32_t,
                                                uint32_t &c, uint32_t &s)
    {
        const uint2 e = reinterpret_cast<const uint2 *>(tab)[sym];
        c = e.x;
        s = e.y;
    }
    __device__ static __forceinline__ uint32_t add_len(uint32_t acc, uint32_t s) { return acc + s; }
};

// right-aligned codeword value from an entry (cold path only)
__device__ __forceinline__ uint32_t entry_cw(uint32_t c, uint32_t s)
{
    const uint32_t len = s & 31u;
    return len ? (c >> (32u - len)) : 0u;
}

// ---- pass 1: look up the thread's 32 symbols, sum their lengths ----------------------------------
template <bool WIDE, bool GUARD>
__device__ __forceinline__ uint32_t lookup32(const uint32_t *tab, const uint32_t (&w)[8],
                                             uint32_t lane, uint32_t nvalid_words,
                                             uint32_t (&c)[kSymPerThread],
                                             uint32_t (&s)[kSymPerThread])
{
    uint32_t bits = 0;
#pragma unroll
    for (int wi = 0; wi < 8; wi++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // cpuencode.cpp:28 -- the most significant byte of a word is its first symbol
            const uint32_t sym = (w[wi] >> (8 * (3 - i))) & 0xFFu;
            uint32_t cc, ss;
            Tab<WIDE>::look(tab, sym, lane, cc, ss);
            if (GUARD && (uint32_t)wi >= nvalid_words) {
                cc = 0;
                ss = 0;
            }
            c[4 * wi + i] = cc;
            s[4 * wi + i] = ss;
            bits = Tab<WIDE>::add_len(bits, ss);
        }
    }
    return bits;
}

// ---- pass 2: append the 32 codewords at tile-relative bit offset q0, emit completed words ---------
// Returns the thread's partial tail word (left-aligned, zero padded; 0 if it ends word-aligned).
template <int G, bool WIDE, bool ATOMIC>
__device__ __forceinline__ uint32_t emit32(const uint32_t (&c)[kSymPerThread],
                                           const uint32_t (&s)[kSymPerThread], uint32_t q0,
                                           uint32_t *stage)
{
    uint32_t hi = 0, lo = 0, q = q0, qflushed = q0;
#pragma unroll
    for (int i = 0; i < kSymPerThread; i++) {
        // (hi:lo) = ((hi:lo) << len) | cw, with len = s & 31 and cw = top `len` bits of c
        hi = __funnelshift_l(lo, hi, s[i]);
        lo = __funnelshift_l(c[i], lo, s[i]);
        q = Tab<WIDE>::add_len(q, s[i]);
        if ((i % G) == G - 1 || i == kSymPerThread - 1) {
            // at most one word boundary can have been crossed since the last check
            if ((q ^ qflushed) & ~31u) {
                const uint32_t word = __funnelshift_r(lo, hi, q);   // the 32 bits above the pending q%32
                if (ATOMIC)
                    atomicOr(&stage[(q >> 5) - 1], word);
                else
                    stage[(q >> 5) - 1] = word;
            }
            qflushed = q;
        }
    }
    const uint32_t f = q & 31u;
    return f ? (lo << (32u - f)) : 0u;
}

// ---- the last `need` (< 32) stream bits that precede symbol index `first_sym` ---------------------
// Executed by one full warp.  Walks backwards 32 symbols at a time until `need` bits are covered or
// the buffer start is reached (then the missing high bits are zero: the start_bit phase of a shard).
template <bool WIDE>
__device__ uint32_t bits_before(const EncParams &p, const uint32_t *tab, unsigned long long first_sym,
                                uint32_t need, uint32_t lane)
{
    const unsigned char *bytes = reinterpret_cast<const unsigned char *>(p.in);
    uint32_t acc = 0, have = 0;
    long long base = (long long)first_sym - 1;
    while (have < need && base >= 0) {
        const long long idx = base - (long long)lane;
        uint32_t cw = 0, len = 0;
        if (idx >= 0) {
            const unsigned long long a = ((unsigned long long)idx & ~3ULL) + (3ULL - ((unsigned long long)idx & 3ULL));
            uint32_t c, s;
            Tab<WIDE>::look(tab, bytes[a], lane, c, s);
            len = s & 31u;
            cw = entry_cw(c, s);
        }
        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += n;
        }
        const uint32_t pos = have + incl - len;            // bit index (from the LSB) of this codeword's LSB
        const uint32_t contrib = (pos < 32u) ? (cw << pos) : 0u;
        acc |= __reduce_or_sync(0xFFFFFFFFu, contrib);
        have += __shfl_sync(0xFFFFFFFFu, incl, 31);
        base -= 32;
    }
    return acc;
}

// ---- the kernel -----------------------------------------------------------------------------------
template <int G, bool WIDE>
__global__ void __launch_bounds__(kEncThreads, 2) encode_kernel(const EncParams p)
{
    using T = Tab<WIDE>;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *tab = smem;
    uint32_t *stage = smem + T::kWords;

    __shared__ unsigned long long s_tile;
    __shared__ unsigned long long s_prefix;
    __shared__ uint32_t s_prev;
    __shared__ uint32_t s_wsum[kWarps];
    __shared__ uint32_t s_wtail[kWarps];

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    T::fill(tab, p.table, tid);
    if (tid == 0)
        s_tile = atomicAdd(p.ticket, 1ULL) - p.ticket_base + p.first_tile;
    __syncthreads();

    for (;;) {
        const unsigned long long tile = s_tile;
        if (tile >= p.end_tile)
            break;

        // ---------------- load + pass 1 ----------------
        const unsigned long long word0 = tile * (unsigned long long)kTileWords;
        const unsigned long long left = p.n_words - word0;
        const bool full = left >= (unsigned long long)kTileWords;
        uint32_t w[8];
        uint32_t c[kSymPerThread], s[kSymPerThread];
        uint32_t bt;
        if (full) {
            ld_stream_v8(p.in + word0 + tid * 8u, w);
            bt = lookup32<WIDE, false>(tab, w, lane, 8, c, s);
        } else {
            const uint32_t mine = tid * 8u;
            const uint32_t nvalid = (left > mine) ? (uint32_t)min((unsigned long long)8, left - mine) : 0u;
#pragma unroll
            for (int i = 0; i < 8; i++)
                w[i] = ((uint32_t)i < nvalid) ? __ldg(p.in + word0 + mine + i) : 0u;
            bt = lookup32<WIDE, true>(tab, w, lane, nvalid, c, s);
        }

        // ---------------- tile-wide exclusive scan of per-thread bit counts ----------------
        uint32_t incl = bt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += n;
        }
        if (lane == 31) s_wsum[warp] = incl;
        // a thread with < 32 bits breaks the "a staging word has at most two owners" rule
        const int slow = __syncthreads_or(bt < 32u);
        uint32_t wsum = (lane < (uint32_t)kWarps) ? s_wsum[lane] : 0u;
        uint32_t wincl = wsum;
#pragma unroll
        for (int d = 1; d < kWarps; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wincl, d);
            if (lane >= (uint32_t)d) wincl += n;
        }
        const uint32_t btile = __shfl_sync(0xFFFFFFFFu, wincl, kWarps - 1);
        const uint32_t wbase = __shfl_sync(0xFFFFFFFFu, wincl - wsum, warp);
        const uint32_t q0 = wbase + incl - bt;               // tile-relative bit offset of this thread

        unsigned long long next_ticket = 0;
        if (tid == 0) {
            // publish early: successors only need the count, not our data
            if (tile == 0)
                st_relaxed_u64(&p.desc[0], pack_desc(p.epoch, kStatusPrefix, p.start_bit + btile));
            else
                st_relaxed_u64(&p.desc[tile], pack_desc(p.epoch, kStatusAggregate, btile));
            next_ticket = atomicAdd(p.ticket, 1ULL);          // latency hidden behind pass 2
        }

        // ---------------- pass 2: bits -> shared staging (tile-relative alignment) ----------------
        const uint32_t nstage = (btile + 31u) >> 5;
        if (!slow) {
            const uint32_t tail = emit32<G, WIDE, false>(c, s, q0, stage);
            const uint32_t left_tail = __shfl_up_sync(0xFFFFFFFFu, tail, 1);
            if (lane == 31) s_wtail[warp] = tail;
            if (lane != 0 && (q0 & 31u)) stage[q0 >> 5] |= left_tail;   // my head word, completed by me
            if (tid == kEncThreads - 1 && (btile & 31u)) stage[btile >> 5] = tail;
            __syncthreads();
            if (lane == 0 && warp != 0 && (q0 & 31u)) stage[q0 >> 5] |= s_wtail[warp - 1];
        } else {
            for (uint32_t j = tid; j <= nstage; j += kEncThreads) stage[j] = 0u;
            __syncthreads();
            const uint32_t tail = emit32<G, WIDE, true>(c, s, q0, stage);
            if (tail) atomicOr(&stage[(q0 + bt) >> 5], tail);
            __syncthreads();
        }

        // ---------------- decoupled look-back (warp 0) ----------------
        if (warp == 0) {
            unsigned long long excl;
            if (tile == 0) {
                excl = p.start_bit;
            } else {
                excl = 0;
                long long look = (long long)tile - 1;
                for (;;) {
                    const long long idx = look - (long long)lane;
                    const unsigned long long d =
                        (idx >= 0) ? ld_relaxed_u64(&p.desc[idx]) : pack_desc(p.epoch, kStatusPrefix, 0);
                    const uint32_t st =
                        ((uint32_t)(d >> 50) == p.epoch) ? (uint32_t)((d >> kDescValueBits) & 3u) : 0u;
                    const uint32_t pmask = __ballot_sync(0xFFFFFFFFu, st == kStatusPrefix);
                    const uint32_t xmask = __ballot_sync(0xFFFFFFFFu, st == 0u);
                    const uint32_t first_p = pmask ? (uint32_t)(__ffs(pmask) - 1) : 32u;
                    const uint32_t need = (first_p >= 31u) ? 0xFFFFFFFFu : ((2u << first_p) - 1u);
                    if (xmask & need) {
                        __nanosleep(32);
                        continue;                               // a needed predecessor has not published yet
                    }
                    unsigned long long v = ((need >> lane) & 1u) ? (d & kDescValueMask) : 0ULL;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                    excl += v;
                    if (first_p < 32u) break;
                    look -= 32;
                }
                if (lane == 0)
                    st_relaxed_u64(&p.desc[tile], pack_desc(p.epoch, kStatusPrefix, excl + btile));
            }
            const uint32_t sh = (uint32_t)(excl & 31ULL);
            uint32_t prev = 0;
            // nothing precedes the job's first bit: the start_bit phase is zero-filled (also keeps
            // all-zero-length codebooks from walking the whole input backwards)
            if (sh != 0 && tile != 0 && excl != p.start_bit)
                prev = bits_before<WIDE>(p, tab, tile * (unsigned long long)kTileBytes, sh, lane);
            if (lane == 0) {
                s_prefix = excl;
                s_prev = prev;
            }
        }
        __syncthreads();

        // ---------------- copy-out: staging -> global, shifted to the global phase ----------------
        {
            const unsigned long long P = s_prefix;
            const uint32_t sh = (uint32_t)(P & 31ULL);
            const unsigned long long g0 = P >> 5;
            const unsigned long long end = P + btile;
            const uint32_t nfull = (uint32_t)((end >> 5) - g0);
            const bool last = (tile == p.n_tiles - 1);
            const uint32_t nwrite = nfull + (last ? 1u : 0u);
            const uint32_t prev = s_prev;
            bool spill = false;
            for (uint32_t j = tid; j < nwrite; j += kEncThreads) {
                const uint32_t cur = (j < nstage) ? stage[j] : 0u;
                const uint32_t before = (j == 0) ? prev : ((j - 1 < nstage) ? stage[j - 1] : 0u);
                const uint32_t v = __funnelshift_r(cur, before, sh);
                if (g0 + j < p.out_cap_words)
                    p.out[g0 + j] = v;
                else if (!(last && j == nfull && (end & 31ULL) == 0))  // the courtesy zero word may not fit
                    spill = true;
            }
            if (spill) p.result->overflow = 1ULL;
            if (tid == 0 && tile == p.end_tile - 1) p.result->bits_end = end;
        }

        if (tid == 0) s_tile = next_ticket - p.ticket_base + p.first_tile;
        __syncthreads();
    }
}

template <int G, bool WIDE>
cudaError_t launch_one(const EncParams &p, int grid, cudaStream_t stream)
{
    const size_t smem = (size_t)(Tab<WIDE>::kWords + Tab<WIDE>::kStageWords) * sizeof(uint32_t);
    encode_kernel<G, WIDE><<<grid, kEncThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int G, bool WIDE>
cudaError_t configure_one()
{
    const size_t smem = (size_t)(Tab<WIDE>::kWords + Tab<WIDE>::kStageWords) * sizeof(uint32_t);
    return cudaFuncSetAttribute(encode_kernel<G, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem);
}

template <int G, bool WIDE>
int occupancy_one()
{
    const size_t smem = (size_t)(Tab<WIDE>::kWords + Tab<WIDE>::kStageWords) * sizeof(uint32_t);
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, encode_kernel<G, WIDE>, kEncThreads, smem) !=
        cudaSuccess)
        return 0;
    return n;
}

}  // namespace

const char *variant_name(EncVariant v)
{
    switch (v) {
    case kPackedG4: return "packed_g4";
    case kPackedG3: return "packed_g3";
    case kPackedG2: return "packed_g2";
    case kPackedG1: return "packed_g1";
    case kWideG1: return "wide_g1";
    default: return "?";
    }
}

EncVariant pick_variant(int max_len)
{
    if (max_len <= 8) return kPackedG4;
    if (max_len <= 10) return kPackedG3;
    if (max_len <= 16) return kPackedG2;
    if (max_len <= kPackedMaxLen) return kPackedG1;
    return kWideG1;
}

size_t encode_smem_bytes(EncVariant v)
{
    return (v == kWideG1) ? (size_t)(kWideTabWords + kWideStageWords) * 4
                          : (size_t)(kPackedTabWords + kPackedStageWords) * 4;
}

cudaError_t encode_configure()
{
    cudaError_t e;
    if ((e = configure_one<4, false>()) != cudaSuccess) return e;
    if ((e = configure_one<3, false>()) != cudaSuccess) return e;
    if ((e = configure_one<2, false>()) != cudaSuccess) return e;
    if ((e = configure_one<1, false>()) != cudaSuccess) return e;
    return configure_one<1, true>();
}

int encode_max_ctas_per_sm(EncVariant v)
{
    switch (v) {
    case kPackedG4: return occupancy_one<4, false>();
    case kPackedG3: return occupancy_one<3, false>();
    case kPackedG2: return occupancy_one<2, false>();
    case kPackedG1: return occupancy_one<1, false>();
    default: return occupancy_one<1, true>();
    }
}

cudaError_t launch_encode(EncVariant v, const EncParams &p, int grid, cudaStream_t stream)
{
    switch (v) {
    case kPackedG4: return launch_one<4, false>(p, grid, stream);
    case kPackedG3: return launch_one<3, false>(p, grid, stream);
    case kPackedG2: return launch_one<2, false>(p, grid, stream);
    case kPackedG1: return launch_one<1, false>(p, grid, stream);
    default: return launch_one<1, true>(p, grid, stream);
    }
}

}  // namespace hb
