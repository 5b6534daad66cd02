/*
 * hb_encode.cu -- single-pass Huffman variable-length encode for sm_100a (B200).
 *
 * Replaces, in ONE kernel, the reference's three passes
 *     vlc_encode_kernel_sm64huff (vlc_kernel_sm64huff.cu:37-160)   per-block encode
 *     prescanArray               (scan.cu:114-231, scanLargeArray_kernel.cu:216-258)   block offsets
 *     cudaMemset + pack2         (main_test_cu.cu:162, pack_kernels.cu:19-52)   compaction
 * and produces the exact word stream of cpu_vlc_encode (cpuencode.cpp:12-46).
 *
 * Design (not a port: none of the reference's structure survives).  The kernel is bound by
 * instruction issue, not by HBM, so everything is arranged to spend as few issue slots per symbol
 * as possible and to never leave a warp waiting on another one:
 *
 *   - one persistent CTA per SM: 16 autonomous WORKER warps + 1 SCOUT warp.  Tiles of 16 KiB are
 *     handed out by an atomic ticket counter; a worker warp owns a 1 KiB chunk of each tile
 *     (32 contiguous symbols per lane, one 256-bit load, prefetched one tile ahead);
 *   - the codebook lives in shared memory at a 256-byte stride per symbol, replicated per lane:
 *     ONE byte-permute builds the whole lookup address (symbol -> byte 1, lane*4 -> byte 0) and
 *     the lookup is bank-conflict free for any symbol distribution.  An entry is
 *     (cw << (32-len)) | len, so ONE funnel shift appends a codeword to a running 32-bit window
 *     (shf.l.wrap takes its shift count from the low 5 bits of the same register) and one dp4a
 *     accumulates the length: 4 issue slots per symbol (prmt, lds, shf, dp4a);
 *   - the window is snapshotted every G symbols.  After a warp shuffle scan has placed the lane
 *     inside the warp's chunk, a second pass over the G-symbol groups only tests "did this group
 *     cross a 32-bit word boundary" and, if so, rebuilds that word from two neighbouring
 *     snapshots with two funnel shifts and stores it to the warp's private staging region.
 *     Because every lane of a full chunk emits >= 32 bits, a staging word has at most two owners:
 *     the left lane hands its partial tail word to the right one through a shuffle -- no
 *     shared-memory atomics, no zeroing.  Chunks that break the rules (ragged end of the input,
 *     zero-length codes, a group of 32+ bits) are re-encoded symbol by symbol with atomicOr;
 *   - workers never synchronise with each other.  They post their chunk's bit count to the scout,
 *     which publishes the tile aggregate and resolves the tile's global bit offset by a decoupled
 *     look-back over 64-bit descriptors {epoch, status, 48-bit count} while the workers are
 *     already encoding the next tile (staging is double buffered; mbarriers carry the hand-offs);
 *   - one tile later each worker copies its own staging region out, coalesced, with one funnel
 *     shift per word to the global phase.  An output word that straddles two chunks belongs to
 *     the right-hand chunk, which takes the missing (< 32) bits from a tiny carry ring; for the
 *     first chunk of a tile the scout re-derives them from the symbols just before the tile, so
 *     there is no inter-CTA data dependency, no atomics on the output and no memset of it.
 */
#include "hb_kernels.cuh"

namespace hb {
namespace {

constexpr int kW = kEncWorkers;
constexpr int S = kSymPerThread;
constexpr int kSlotBytes = 256;                         // table stride per symbol
constexpr int kTabWords = 256 * kSlotBytes / 4;         // 64 KiB
constexpr unsigned long long kNoTile = ~0ULL;

// Shared-memory map.  The table must start on a 64 KiB boundary of the CTA's shared window so that
// the byte permute can produce a complete lookup address (window address bytes 2..3 are constants).
// The window starts with kSmemReserved bytes owned by the system, so the dynamic block is laid out as
//   [staging slot 0 | pad] up to the boundary, [table 64 KiB], [staging slot 1], [control block].
constexpr uint32_t kSmemReserved = 1024;                // cudaDevAttrReservedSharedMemoryPerBlock on sm_100
constexpr uint32_t kTabOffset = 65536 - kSmemReserved;  // table offset inside the dynamic block

template <bool WIDE>
struct Geo {
    // a chunk of 32*S symbols can emit at most 32*S*max_len bits (max_len 24 packed, 31 wide)
    static constexpr int kRegionWords = S * (WIDE ? 31 : 24);
    static constexpr uint32_t kSlotBytes = kW * kRegionWords * 4;
    static constexpr uint32_t kSlot1Offset = kTabOffset + kTabWords * 4;
    static constexpr uint32_t kCtrlOffset = kSlot1Offset + kSlotBytes;
    static_assert(kSlotBytes <= kTabOffset, "staging slot 0 must fit below the table");
};

struct Ctrl {
    unsigned long long bar_sums[2];     // workers -> scout: chunk bit counts of tile k posted
    unsigned long long bar_emit[2];     // workers -> workers: chunk carries of tile k posted
    unsigned long long bar_prefix[2];   // scout -> workers: global offset of tile k resolved
    unsigned long long bar_tile[4];     // scout -> workers: ring[k & 3] holds the k-th tile id
    unsigned long long ring[4];
    unsigned long long prefix[2];
    uint32_t prev[2];
    uint32_t flags[2];
    uint32_t woff[2][kW];
    uint32_t sums[2][kW];
    uint32_t carry_val[4][kW];
    uint32_t carry_cnt[4][kW];
};

// ---- small PTX helpers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    const uint32_t a = smem_addr(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}
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

// ---- codebook in shared memory ---------------------------------------------------------------------
// slot(sym) = 256 bytes: words 0..31 = the entry replicated per lane; wide tables keep the length in
// words 32..63.  Lookups take a complete shared-window address (table base + sym*256 + lane*4).
__device__ __forceinline__ uint32_t tab_ld(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t tab_ld_len(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+128];" : "=r"(v) : "r"(addr));
    return v;
}
template <bool WIDE>
__device__ __forceinline__ void fetch_entry(uint32_t tab_s, uint32_t sym, uint32_t lane, uint32_t &c,
                                            uint32_t &len)
{
    const uint32_t addr = tab_s + sym * kSlotBytes + lane * 4u;
    c = tab_ld(addr);
    len = WIDE ? tab_ld_len(addr) : (c & 0xFFu);
}
template <bool WIDE>
__device__ __forceinline__ void fill_table(uint32_t *tab, const uint32_t *g, uint32_t tid)
{
    // packed source: uint32[256]; wide source: {cw_left, len}[256]
    for (uint32_t i = tid; i < 256u * 64u; i += kEncThreads) {
        const uint32_t sym = i >> 6, k = i & 63u;
        uint32_t v = 0;
        if (WIDE)
            v = __ldg(g + 2u * sym + (k >> 5));
        else if (k < 32u)
            v = __ldg(g + sym);
        tab[i] = v;
    }
}

// symbol `idx` (in encode order) of the job lives at this byte of the little-endian word buffer:
// cpuencode.cpp:28 -- the most significant byte of a word is its first symbol
__device__ __forceinline__ unsigned long long byte_of_symbol(unsigned long long idx)
{
    return (idx & ~3ULL) + (3ULL - (idx & 3ULL));
}

// ---- the last `need` (< 32) stream bits that precede symbol index `first_sym` ---------------------
// Executed by one full warp.  Walks backwards 32 symbols at a time until `need` bits are covered or
// the buffer start is reached (then the missing high bits are zero: the start_bit phase of a shard).
template <bool WIDE>
__device__ uint32_t bits_before(const EncParams &p, uint32_t tab_s,
                                unsigned long long first_sym, uint32_t need, uint32_t lane)
{
    const unsigned char *bytes = reinterpret_cast<const unsigned char *>(p.in);
    uint32_t acc = 0, have = 0;
    long long base = (long long)first_sym - 1;
    while (have < need && base >= 0) {
        const long long idx = base - (long long)lane;
        uint32_t cw = 0, len = 0;
        if (idx >= 0) {
            uint32_t c;
            fetch_entry<WIDE>(tab_s, bytes[byte_of_symbol((unsigned long long)idx)], lane, c, len);
            cw = len ? (c >> (32u - len)) : 0u;
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

// ---- scout warp: tickets, aggregates, look-back -------------------------------------------------------
template <bool WIDE>
__device__ void scout(const EncParams &p, uint32_t tab_s, Ctrl *ctrl, uint32_t lane)
{
    bool ended = false;
    // k-th tile of this CTA -> ring[k & 3]
    auto post = [&](uint32_t k) -> unsigned long long {
        unsigned long long t = kNoTile;
        if (!ended) {
            unsigned long long tk = 0;
            if (lane == 0) tk = atomicAdd(p.ticket, 1ULL);
            tk = __shfl_sync(0xFFFFFFFFu, tk, 0);
            t = tk - p.ticket_base + p.first_tile;
            if (t >= p.end_tile) {
                t = kNoTile;
                ended = true;           // exactly one ticket past the end per CTA
            }
        }
        if (lane == 0) {
            ctrl->ring[k & 3u] = t;
            mbar_arrive(&ctrl->bar_tile[k & 3u]);
        }
        return t;
    };

    unsigned long long t_cur = post(0);
    unsigned long long t_next = post(1);
    for (uint32_t k = 0; t_cur != kNoTile; k++) {
        const uint32_t slot = k & 1u;
        const unsigned long long tile = t_cur;
        mbar_wait(&ctrl->bar_sums[slot], (k >> 1) & 1u);

        // exclusive scan of the 16 chunk bit counts
        const uint32_t n = (lane < (uint32_t)kW) ? ctrl->sums[slot][lane] : 0u;
        uint32_t incl = n;
#pragma unroll
        for (int d = 1; d < kW; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += v;
        }
        const uint32_t btile = __shfl_sync(0xFFFFFFFFu, incl, kW - 1);
        if (lane < (uint32_t)kW) ctrl->woff[slot][lane] = incl - n;

        // publish early: successors only need the count, not our data
        if (lane == 0) {
            if (tile == 0)
                st_relaxed_u64(&p.desc[0], pack_desc(p.epoch, kStatusPrefix, p.start_bit + btile));
            else
                st_relaxed_u64(&p.desc[tile], pack_desc(p.epoch, kStatusAggregate, btile));
        }
        // every worker is past pass 1 of tile k, so ring[(k + 2) & 3] (tile k - 2) is dead
        t_cur = t_next;
        t_next = post(k + 2);

        // ---------------- decoupled look-back ----------------
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
                    __nanosleep(20);
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
            prev = bits_before<WIDE>(p, tab_s, tile * (unsigned long long)kTileBytes, sh, lane);
        if (lane == 0) {
            ctrl->prefix[slot] = excl;
            ctrl->prev[slot] = prev;
            ctrl->flags[slot] = (tile == p.n_tiles - 1 ? 1u : 0u) | (tile == p.end_tile - 1 ? 2u : 0u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->bar_prefix[slot]);
    }
}

// ---- worker: copy one staged chunk to its place in the global stream ---------------------------------
__device__ __forceinline__ void copy_out(const EncParams &p, Ctrl *ctrl, const uint32_t *st, uint32_t k,
                                         uint32_t n, uint32_t warp, uint32_t lane)
{
    const uint32_t slot = k & 1u;
    mbar_wait(&ctrl->bar_emit[slot], (k >> 1) & 1u);
    mbar_wait(&ctrl->bar_prefix[slot], (k >> 1) & 1u);

    // the (< 32) bits that precede this chunk: neighbours' carries, then the tile's `prev`
    uint32_t cin = 0, have = 0;
    for (int r = (int)warp - 1; r >= 0 && have < 31u; r--) {
        cin |= ctrl->carry_val[k & 3u][r] << have;
        have += ctrl->carry_cnt[k & 3u][r];
    }
    if (have < 31u) cin |= ctrl->prev[slot] << have;

    const unsigned long long B = ctrl->prefix[slot] + ctrl->woff[slot][warp];
    const uint32_t flags = ctrl->flags[slot];
    const uint32_t sh = (uint32_t)(B & 31ULL);
    const unsigned long long g0 = B >> 5;
    const unsigned long long end = B + n;
    const uint32_t nfull = (uint32_t)((end >> 5) - g0);
    const bool last = (flags & 1u) && warp == (uint32_t)kW - 1;   // the job's final word(s)
    const uint32_t nwrite = nfull + (last ? 1u : 0u);
    const uint32_t nstage = (n + 31u) >> 5;
    bool spill = false;
    for (uint32_t j = lane; j < nwrite; j += 32u) {
        const uint32_t cur = (j < nstage) ? st[j] : 0u;
        const uint32_t before = (j == 0) ? cin : ((j - 1 < nstage) ? st[j - 1] : 0u);
        const uint32_t v = __funnelshift_r(cur, before, sh);
        if (g0 + j < p.out_cap_words)
            p.out[g0 + j] = v;
        else if (!(last && j == nfull && (end & 31ULL) == 0))      // the courtesy zero word may not fit
            spill = true;
    }
    if (spill) p.result->overflow = 1ULL;
    if ((flags & 2u) && warp == (uint32_t)kW - 1 && lane == 0) p.result->bits_end = end;
}

// ---- worker warp ----------------------------------------------------------------------------------------
template <int G, bool WIDE, bool CHECK>
__device__ void worker(const EncParams &p, uint32_t tab_s, uint32_t *stage0, uint32_t *stage1, Ctrl *ctrl,
                       uint32_t warp, uint32_t lane)
{
    constexpr int NG = (S + G - 1) / G;
    constexpr int RW = Geo<WIDE>::kRegionWords;
    // byte 0 = lane*4, bytes 1..2 = bytes 2..3 of the table's window address (prmt source b)
    const uint32_t laneoff = lane * 4u | ((tab_s >> 16) << 8);
    const unsigned char *bytes = reinterpret_cast<const unsigned char *>(p.in);
    const unsigned long long n_bytes = p.n_words * 4ULL;

    auto region = [&](uint32_t slot) { return (slot ? stage1 : stage0) + warp * RW; };

    uint32_t w[8], wn[8];
    mbar_wait(&ctrl->bar_tile[0], 0);
    unsigned long long tile = ctrl->ring[0];
    auto chunk_word0 = [&](unsigned long long t) {
        return t * (unsigned long long)kTileWords + warp * (unsigned long long)(kChunkBytes / 4);
    };
    auto chunk_full = [&](unsigned long long t) {
        return chunk_word0(t) + (unsigned long long)(kChunkBytes / 4) <= p.n_words;
    };
    if (tile != kNoTile && chunk_full(tile)) ld_stream_v8(p.in + chunk_word0(tile) + lane * 8u, w);

    uint32_t n_prev = 0;
    uint32_t k = 0;
    for (; tile != kNoTile; k++) {
        const uint32_t slot = k & 1u;
        uint32_t *st = region(slot);

        // ---------------- prefetch the next tile's chunk ----------------
        mbar_wait(&ctrl->bar_tile[(k + 1) & 3u], ((k + 1) >> 2) & 1u);
        const unsigned long long tnext = ctrl->ring[(k + 1) & 3u];
        if (tnext != kNoTile && chunk_full(tnext)) ld_stream_v8(p.in + chunk_word0(tnext) + lane * 8u, wn);

        // ---------------- pass 1: look up, chain codewords, sum lengths ----------------
        const bool full = chunk_full(tile);                   // warp-uniform
        const unsigned long long sym0 =
            tile * (unsigned long long)kTileBytes + warp * (unsigned long long)kChunkBytes + lane * (unsigned long long)S;
        uint32_t los[NG], gss[NG];
        uint32_t bt = 0, ormask = 0;
        if (full) {
            uint32_t lo = 0, gs = 0;
#pragma unroll
            for (int i = 0; i < S; i++) {
                // {lane*4, symbol, table address bytes 2..3}: the whole lookup address in one prmt
                const uint32_t off = __byte_perm(w[i >> 2], laneoff, 0x6504u | ((3u - (i & 3)) << 4));
                if (WIDE) {
                    const uint32_t c = tab_ld(off);
                    const uint32_t l = tab_ld_len(off);
                    lo = __funnelshift_l(c, lo, l);
                    gs += l;
                } else {
                    const uint32_t e = tab_ld(off);
                    lo = __funnelshift_l(e, lo, e);           // (lo << len) | cw, len = e & 31
                    gs = __dp4a(e, 1u, gs);                   // + (e & 0xFF)
                }
                if ((i % G) == G - 1 || i == S - 1) {
                    los[i / G] = lo;
                    gss[i / G] = gs;
                    bt += gs;
                    if (CHECK) ormask |= gs;
                    gs = 0;
                }
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < S; i++) {
                if (sym0 + i < n_bytes) {
                    uint32_t c, l;
                    fetch_entry<WIDE>(tab_s, bytes[byte_of_symbol(sym0 + i)], lane, c, l);
                    bt += l;
                }
            }
#pragma unroll
            for (int g = 0; g < NG; g++) los[g] = gss[g] = 0;
        }

        // ---------------- warp scan: this lane's bit offset inside the chunk ----------------
        uint32_t incl = bt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += v;
        }
        const uint32_t q0 = incl - bt;
        const uint32_t n = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (lane == 31) {
            ctrl->sums[slot][warp] = n;
            mbar_arrive(&ctrl->bar_sums[slot]);
        }

        // ---------------- pass 2: bits -> this warp's staging region (chunk-relative alignment) ----------------
        // fast path: a staging word has at most two owners (needs >= 32 bits from every lane) and
        // every group fits the 32-bit window
        const bool fast = full && __all_sync(0xFFFFFFFFu, bt >= 32u && (!CHECK || (ormask & ~31u) == 0u));
        if (fast) {
            uint32_t q = q0;
            uint32_t *wp = st + (q0 >> 5);                    // the word this lane completes next
            uint32_t lo_prev = 0;
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const uint32_t qn = q + gss[g];
                if ((qn ^ q) & ~31u) {
                    // the 32 bits that end at the boundary: low (qn & 31) of them come from the window
                    // before this group, the rest from the window after it
                    const uint32_t hi = __funnelshift_l(lo_prev, 0u, gss[g]);   // lo_prev >> (32 - gs)
                    *wp++ = __funnelshift_r(los[g], hi, qn);
                }
                q = qn;
                lo_prev = los[g];
            }
            const uint32_t f = q & 31u;
            const uint32_t tail = f ? (lo_prev << (32u - f)) : 0u;
            const uint32_t left_tail = __shfl_up_sync(0xFFFFFFFFu, tail, 1);
            if (lane != 0 && (q0 & 31u)) st[q0 >> 5] |= left_tail;   // my head word, completed by me
            if (lane == 31 && f) st[n >> 5] = tail;
        } else {
            for (uint32_t j = lane; j < ((n + 31u) >> 5); j += 32u) st[j] = 0u;
            __syncwarp();
            uint32_t q = q0, lo = 0;
#pragma unroll 1
            for (int i = 0; i < S; i++) {
                if (sym0 + i < n_bytes) {
                    uint32_t c, l;
                    fetch_entry<WIDE>(tab_s, bytes[byte_of_symbol(sym0 + i)], lane, c, l);
                    if (l) {
                        const uint32_t ln = __funnelshift_l(c, lo, l);
                        const uint32_t qn = q + l;
                        if ((qn ^ q) & ~31u)
                            atomicOr(&st[(qn >> 5) - 1u], __funnelshift_r(ln, __funnelshift_l(lo, 0u, l), qn));
                        q = qn;
                        lo = ln;
                    }
                }
            }
            const uint32_t f = q & 31u;
            if (f) atomicOr(&st[q >> 5], lo << (32u - f));
        }
        __syncwarp();

        // ---------------- carry: the last (<= 31) bits of this chunk, for the right-hand neighbour ----------------
        if (lane == 0) {
            uint32_t val = 0;
            if (n) {
                const uint32_t a = (n - 1u) >> 5, r = n & 31u;
                const uint32_t w1 = st[a], w0 = a ? st[a - 1u] : 0u;
                val = (r ? __funnelshift_l(w1, w0, r) : w1) & 0x7FFFFFFFu;
            }
            ctrl->carry_val[k & 3u][warp] = val;
            ctrl->carry_cnt[k & 3u][warp] = n < 31u ? n : 31u;
            mbar_arrive(&ctrl->bar_emit[slot]);
        }

        // ---------------- copy-out of the PREVIOUS tile (its look-back had a whole tile of slack) ----------------
        if (k > 0) copy_out(p, ctrl, region((k - 1u) & 1u), k - 1u, n_prev, warp, lane);

        __syncwarp();
        n_prev = n;
        tile = tnext;
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = wn[i];
    }
    if (k > 0) copy_out(p, ctrl, region((k - 1u) & 1u), k - 1u, n_prev, warp, lane);
}

// ---- the kernel -----------------------------------------------------------------------------------
template <int G, bool WIDE, bool CHECK>
__global__ void __launch_bounds__(kEncThreads, 1) encode_kernel(const EncParams p)
{
    extern __shared__ __align__(1024) uint32_t smem[];
    unsigned char *base = reinterpret_cast<unsigned char *>(smem);
    uint32_t *tab = reinterpret_cast<uint32_t *>(base + kTabOffset);
    uint32_t *stage0 = smem;
    uint32_t *stage1 = reinterpret_cast<uint32_t *>(base + Geo<WIDE>::kSlot1Offset);
    Ctrl *ctrl = reinterpret_cast<Ctrl *>(base + Geo<WIDE>::kCtrlOffset);
    const uint32_t tab_s = smem_addr(tab);
    if (tab_s & 0xFFFFu) {
        // the shared window is not laid out as assumed: refuse loudly instead of mis-encoding
        if (threadIdx.x == 0) p.result->overflow = 2ULL;
        return;
    }

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    fill_table<WIDE>(tab, p.table, tid);
    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&ctrl->bar_sums[i], kW);
            mbar_init(&ctrl->bar_emit[i], kW);
            mbar_init(&ctrl->bar_prefix[i], 1);
        }
        for (int i = 0; i < 4; i++) mbar_init(&ctrl->bar_tile[i], 1);
    }
    __syncthreads();

    if (warp == (uint32_t)kW)
        scout<WIDE>(p, tab_s, ctrl, lane);
    else
        worker<G, WIDE, CHECK>(p, tab_s, stage0, stage1, ctrl, warp, lane);
}

template <bool WIDE>
constexpr size_t smem_bytes()
{
    return (size_t)Geo<WIDE>::kCtrlOffset + sizeof(Ctrl);
}

// ---- variant table ------------------------------------------------------------------------------------
typedef void (*KernelFn)(const EncParams);
struct VariantRow {
    int group;
    bool wide, check;
    KernelFn fn;
    size_t smem;
    const char *name;
};
#define HB_ROW(G, W, C, NAME) {G, W, C, encode_kernel<G, W, C>, smem_bytes<W>(), NAME}
const VariantRow kRows[] = {
    HB_ROW(8, false, false, "packed_g8"),  HB_ROW(8, false, true, "packed_g8c"),
    HB_ROW(6, false, false, "packed_g6"),  HB_ROW(6, false, true, "packed_g6c"),
    HB_ROW(4, false, false, "packed_g4"),  HB_ROW(4, false, true, "packed_g4c"),
    HB_ROW(3, false, false, "packed_g3"),  HB_ROW(3, false, true, "packed_g3c"),
    HB_ROW(2, false, false, "packed_g2"),  HB_ROW(2, false, true, "packed_g2c"),
    HB_ROW(1, false, false, "packed_g1"),
    HB_ROW(4, true, true, "wide_g4c"),     HB_ROW(2, true, true, "wide_g2c"),
    HB_ROW(1, true, false, "wide_g1"),
};
#undef HB_ROW
constexpr int kNumRows = (int)(sizeof(kRows) / sizeof(kRows[0]));

const VariantRow *find_row(const EncVariant &v)
{
    for (int i = 0; i < kNumRows; i++)
        if (kRows[i].group == v.group && kRows[i].wide == v.wide && kRows[i].check == v.check) return &kRows[i];
    return nullptr;
}

}  // namespace

const char *variant_name(const EncVariant &v)
{
    const VariantRow *r = find_row(v);
    return r ? r->name : "?";
}

EncVariant pick_variant(const uint32_t lens[256])
{
    // implied symbol probabilities 2^-len, normalised (arbitrary tables need not satisfy Kraft)
    double pl[32] = {0};
    double total = 0;
    int max_len = 0;
    for (int s = 0; s < 256; s++) {
        const int l = (int)lens[s];
        if (l <= 0 || l > 31) continue;
        const double pr = 1.0 / (double)(1ULL << l);
        pl[l] += pr;
        total += pr;
        if (l > max_len) max_len = l;
    }
    EncVariant v;
    v.wide = max_len > 24;
    v.group = 1;
    v.check = false;
    if (total <= 0) return v;
    for (int l = 0; l < 32; l++) pl[l] /= total;

    static const int packed_groups[] = {8, 6, 4, 3, 2};
    static const int wide_groups[] = {4, 2};
    const int *cand = v.wide ? wide_groups : packed_groups;
    const int ncand = v.wide ? 2 : 5;
    // a warp falls back to the symbol-by-symbol path when any of its 32*ceil(S/G) groups is >= 32 bits;
    // keep that below ~1% of the chunks
    for (int ci = 0; ci < ncand; ci++) {
        const int G = cand[ci];
        if (G * max_len <= 31) {
            if (v.wide) break;                       // (never true for wide tables)
            v.group = G;
            v.check = false;
            return v;
        }
        // distribution of the sum of G lengths, capped at 32
        double dist[33] = {0}, next[33];
        dist[0] = 1.0;
        for (int j = 0; j < G; j++) {
            for (int x = 0; x <= 32; x++) next[x] = 0;
            for (int x = 0; x <= 32; x++) {
                if (dist[x] == 0) continue;
                for (int l = 1; l < 32; l++) {
                    if (pl[l] == 0) continue;
                    const int y = (x + l > 32) ? 32 : x + l;
                    next[y] += dist[x] * pl[l];
                }
            }
            for (int x = 0; x <= 32; x++) dist[x] = next[x];
        }
        const double p_group = dist[32];
        const double groups_per_chunk = 32.0 * (double)((S + G - 1) / G);
        if (p_group * groups_per_chunk <= 0.01) {
            v.group = G;
            v.check = true;
            return v;
        }
    }
    return v;                                        // G = 1: a single codeword always fits
}

size_t encode_smem_bytes(const EncVariant &v) { return v.wide ? smem_bytes<true>() : smem_bytes<false>(); }

cudaError_t encode_configure()
{
    for (int i = 0; i < kNumRows; i++) {
        const cudaError_t e = cudaFuncSetAttribute(kRows[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)kRows[i].smem);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_encode(const EncVariant &v, const EncParams &p, int grid, cudaStream_t stream)
{
    const VariantRow *r = find_row(v);
    if (!r) return cudaErrorInvalidValue;
    r->fn<<<grid, kEncThreads, r->smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace hb
