// bsq_device.cuh -- sm_100a kernels of the FASTQ hot path.
//
// Two streaming passes over a window (<= 2 GiB of the byte stream, 16-byte aligned base):
//
//   k_summarize   every CTA owns a contiguous RUN of 16 KiB tiles.  Tiles arrive in shared memory
//                 as TMA bulk copies (cp.async.bulk + mbarrier).  Per tile: 16-byte shared loads ->
//                 byte-lane compare -> newline bitmap -> block prefix scan.  The run is reduced to
//                 a 64-byte BsqSummary.
//   k_scan_runs   one CTA scans the run summaries (tile_math.h) and gives every run the state it
//                 starts from (newline rank, previous newline positions, SoA destinations).
//   k_resolve     same tiling; with the prefix known every line that ENDS in a tile is resolved
//                 in place: '@' / '+' / length checks (utils.mojo:448-462), id strip
//                 (utils.mojo:221-242), ASCII and quality-range validation from the HI/BAD bitmaps
//                 (record.mojo:76-116), the line-end table for views() and the FastqBatch SoA
//                 (record_batch.mojo:77-87): the lines are staged in shared memory in destination
//                 layout and written back with TMA bulk stores.
//
// No CTA ever waits on another CTA: the only cross-CTA dependency is the kernel boundary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tile_math.h"

namespace bsq {

#ifndef BSQ_TILE
#define BSQ_TILE 16384
#endif
constexpr int kTile = BSQ_TILE;           // bytes per tile
#ifndef BSQ_THREADS
#define BSQ_THREADS 128
#endif
constexpr int kThreads = BSQ_THREADS;     // threads per CTA
constexpr int kWarps = kThreads / 32;
#ifndef BSQ_STAGES
#define BSQ_STAGES 1
#endif
#ifndef BSQ_COPY_STAGED
#define BSQ_COPY_STAGED 1
#endif
#ifndef BSQ_RESOLVE_CTAS
#define BSQ_RESOLVE_CTAS (BSQ_COPY_STAGED ? 5 : 6)
#endif
constexpr int kStages = BSQ_STAGES;       // TMA ring depth per CTA
constexpr int kResolveCtas = BSQ_RESOLVE_CTAS;  // resident CTAs per SM, k_resolve (shared memory + registers)
#ifndef BSQ_VIEW_CTAS
#define BSQ_VIEW_CTAS 6
#endif
constexpr int kViewCtas = BSQ_VIEW_CTAS;  // ... of the instantiations that do not pack (no stage buffer)
// k_summarize comes in two shapes, chosen so that a window's runs (a CTA's static share) fill whole waves:
//   packing passes (kSums): runs per SM = 2 x kResolveCtas = 10 -> one tile buffer (the next tile waits in L2),
//                           10 CTAs / SM: ONE wave
//   the others:             runs per SM = 2 x kViewCtas = 12    -> two tile buffers, 6 CTAs / SM: two waves
template <bool kSums> struct SumShape {
    static constexpr int kStagesOf = kSums ? 1 : 2;
    static constexpr int kCtas = kSums ? 2 * BSQ_RESOLVE_CTAS : BSQ_VIEW_CTAS;
};
// runs per SM and window: a whole number of waves for the kernel that dominates the pass (a run is a
// CTA's static share, so a partial last wave costs a full one): packing passes follow k_resolve (pack),
// the others k_summarize / k_resolve (views)
constexpr int kRunsPerSmPack = 2 * kResolveCtas;
constexpr int kRunsPerSmView = 2 * kViewCtas;
constexpr int kChunks = kTile / 16;       // 16-byte chunks per tile (1024)
constexpr int kChunksPerThread = kChunks / kThreads;  // 8
constexpr int kWords = kTile / 32;        // bitmap words per tile (512)
constexpr int kWordsPerThread = kWords / kThreads;    // 4 -> a thread ranks 128 contiguous bytes
#ifndef BSQ_NLCAP
#define BSQ_NLCAP 512
#endif
constexpr int kNlCap = BSQ_NLCAP;         // newline-list capacity per pass over a tile (reads shorter than
                                          // ~60 bp fill a 16 KiB tile with more newlines: several passes)
constexpr int kHead = 4;                  // carried newline positions in front of the list
constexpr int kLinesCap = kNlCap / 4 + 3; // lines of one class per pass (+ two sentinels)
constexpr int kTilePad = 32;              // readable slack after a tile for unaligned 16-byte loads
constexpr int kHalo = 1024;               // bytes before the tile kept in shared memory too: a line that began
                                          // up to kHalo bytes before the tile is still read from shared memory
// SoA copy of k_resolve: 1 = per-line copy into a shared-memory stage laid out like the destination, written
// back with TMA bulk stores; 0 = destination-ordered direct copy to global memory (no stage buffer)
constexpr bool kCopyStaged = BSQ_COPY_STAGED != 0;
#ifndef BSQ_ROT
#define BSQ_ROT 2          // staged copy: lanes whose lines start in the same bank start at different words
                           // (2: the rank comes from the line distance; 1: from a per-warp bank census, REDUX + MATCH.ANY)
#endif
#ifndef BSQ_INTERLEAVE
#define BSQ_INTERLEAVE 1   // staged copy: sequence and quality lines alternate over the lanes
#endif
constexpr int kStage = kHalo + kTile + 128;  // SoA staging: the lines that END in a tile lie in [halo | tile]; + 3 x 32 alignment slack
constexpr int kMaxWindows = 64;

static_assert(kThreads % 4 == 0 && (kWordsPerThread == 4 || kWordsPerThread == 2) && kChunksPerThread * kThreads == kChunks, "");
constexpr uint32_t kBytesPerThread = 32u * kWordsPerThread;   // contiguous bytes a thread ranks

struct NlWords { uint32_t w[kWordsPerThread]; };              // a thread's share of the newline bitmap

struct WinParams {
    const uint8_t* base;     // window base, 16-byte aligned
    uint32_t begin, end;     // valid bytes [begin, end) relative to base
    uint32_t first_tile;     // begin / kTile
    uint32_t n_tiles;        // tiles [first_tile, n_tiles) cover [begin, end)
    uint32_t tiles_per_run;
    uint32_t n_runs;
    // hand-off from k_summarize to k_resolve (may be null: nothing is written / nothing is read):
    uint32_t* nl_count;      // [tile - first_tile] newlines of the tile
    uint16_t* nl_list;       // [tile - first_tile][kNlCap] tile-relative newline positions, valid when count <= kNlCap
};

struct ScanOut {             // written by k_scan_runs, read by the host
    BsqTotals totals;        // 32 B
    BsqSummary end_state;    // 64 B: init (+) all runs
    BsqSummary region;       // 64 B: all runs WITHOUT the window init (for shard stitching)
};

struct ResolveParams {
    const BsqPrefix* run_pre;
    uint32_t n_complete;         // complete records of this window
    uint32_t id_fast;            // 1: ids are packed here on the assumption that none needs stripping;
                                 //    *strip_flag is raised if one does and the host redoes the ids
    uint32_t* strip_flag;
    unsigned long long* bases;   // sum of the sequence lengths of the complete records (all windows)
    uint32_t rec_mod;            // rec_base % batch_size
    int64_t rec_div;             // rec_base / batch_size
    int64_t rec_base;            // arena index of the window's first record
    int64_t first_record;        // global index of the pass's first record (error context)
    // views()
    uint32_t* line_ends;         // [newlines + 1]
    uint32_t* id_spans;          // pass-wide, already offset to this window: [2 * n_complete]
    // batches(): arena pointers are pass-wide; *_base64 = bytes written by earlier windows
    uint8_t* seq_out; uint8_t* qual_out; uint8_t* id_out;
    int64_t seq_base64, qual_base64, id_base64;
    int64_t* ends_abs; int64_t* id_ends_abs;      // pass-wide, indexed by arena record
    int64_t* ends_base; int64_t* id_ends_base;    // per batch: cumulative at the batch start
    int64_t id_cap;              // bytes allocated for id_out
    int32_t batch_size;
    uint32_t lower, upper;       // quality bounds
    uint32_t rec_limit;          // a record longer than this many bytes does not fit the reference's buffer
                                 // (buffer_capacity, or buffer_max_capacity with growth): parser.mojo:484-503
    uint32_t q5_width;           // 0, or the SIMD width W of record.mojo:90-102 to reproduce (see kQual walk)
    unsigned long long* err;     // min over ((global record << 8) | code)
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D TMA bulk copy
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a TMA that never lands (bad address) must fault the kernel, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// shared -> global bulk copy (bulk async-group); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the bulk stores issued by this thread have finished READING shared memory (the source may be reused)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------------------------------------
// shared-memory layout of one CTA
// ------------------------------------------------------------------------------------------------

struct alignas(128) TileSmem {
    alignas(128) uint8_t data[kStages][kHalo + kTile + kTilePad];  // TMA destinations: [halo | tile | pad]
    alignas(8) uint64_t full_bar[kStages];
    uint32_t warp_tot[2][kWarps][4];          // block scans, double buffered: one barrier per scan
    uint32_t carry[8];
    uint32_t k1_head[8];                      // k_summarize: [0..3] last four newlines, [4..7] first four
    uint32_t k1_red[kWarps * 4];
    alignas(16) uint32_t bm_nl[kWords];       // 1 bit per byte: '\n'
    // ---- k_resolve only ----
    uint32_t nlx[kHead + kNlCap];             // nlx[kHead + j] = position of local newline j;
                                              // nlx[kHead-1-i] = i-th newline before the list
    uint32_t sdst[3][kLinesCap];              // per class stream: destination of each line
    uint32_t ssrc[3][kLinesCap];              //                   source position of each line
    alignas(16) uint16_t nl16[kStages][kNlCap];   // the tile's newline list from k_summarize (TMA destination)
    uint32_t tile_total[kStages + 1];         // newline count words of the tiles in flight (ring, written one tile ahead)
#if BSQ_COPY_STAGED
    // SoA bytes of the pass, laid out like the destination (mod 16): [id | seq | qual], written back with TMA
    // bulk stores.  The HI / BAD validation bitmaps live at its start: they are consumed before the SoA bytes
    // of the tile are staged.  (Passes that do not pack allocate only the bitmaps.)
    alignas(128) uint8_t stage[kStage];
    __device__ __forceinline__ uint32_t* bm_hi_p() { return reinterpret_cast<uint32_t*>(stage); }            // 1 bit per byte: bit 7 set
    __device__ __forceinline__ uint32_t* bm_bad_p() { return reinterpret_cast<uint32_t*>(stage) + kWords; }  // outside [lower, upper]
#else
    // ---- validating instantiations only (the allocation ends here otherwise) ----
    alignas(16) uint32_t bm_hi[kWords];       // 1 bit per byte: bit 7 set
    uint32_t bm_bad[kWords];                  // 1 bit per byte: outside [lower, upper]
    __device__ __forceinline__ uint32_t* bm_hi_p() { return bm_hi; }
    __device__ __forceinline__ uint32_t* bm_bad_p() { return bm_bad; }
#endif
};
static_assert(2 * kWords * 4 <= kStage, "the validation bitmaps fit the stage buffer");

// k_summarize: the same front end on its own (double-buffered) ring
static_assert(kNlCap * 2 <= kHalo, "k_summarize keeps the newline list in an unused halo");
template <int kSumStages>
struct alignas(128) SumSmemT {
    alignas(128) uint8_t data[kSumStages][kHalo + kTile + kTilePad];
    alignas(8) uint64_t full_bar[kSumStages];
    uint32_t warp_tot[2][kWarps][4];
    uint32_t carry[8];
    uint32_t k1_head[8];                      // [0..3] last four newlines, [4..7] first four
    uint32_t k1_red[kWarps * 4];
    uint32_t dirty;                           // validation screen of the current tile: bit 0 HI, bit 1 BAD
    alignas(16) uint32_t bm_nl[kWords];
    // tile-relative newline positions, in order (handed to k_resolve): they live in the halo bytes of ring
    // slot 0, which k_summarize never loads
    __device__ __forceinline__ uint16_t* list() { return reinterpret_cast<uint16_t*>(data[0]); }
    __device__ __forceinline__ uint32_t* bm_hi_p() { return nullptr; }    // (never used: build_bitmaps<false, false>)
    __device__ __forceinline__ uint32_t* bm_bad_p() { return nullptr; }
};

// ------------------------------------------------------------------------------------------------
// tile pipeline: a CTA walks its run, tile by tile, through a kStages-deep TMA ring
// ------------------------------------------------------------------------------------------------

struct TileCursor {
    uint32_t tile;        // current tile index in the window
    uint32_t lo, hi;      // valid window offsets of the tile [lo, hi)
    uint32_t origin;      // window offset of data[stage][0]
    uint32_t stage;
};

__device__ __forceinline__ uint32_t tile_bytes_rounded(const WinParams& W, uint32_t tile) {
    const uint32_t origin = tile * (uint32_t)kTile;
    uint32_t n = W.end - origin;
    if (n > (uint32_t)kTile) n = kTile;
    return (n + 15u) & ~15u;  // stays inside the 16-byte granule that holds the last valid byte
}

// list_count: newlines of the tile when its ordered list (k_summarize's hand-off) is to arrive with it, else 0
template <bool kWithHalo, typename SM>
__device__ __forceinline__ void issue_tile_load(SM& S, const WinParams& W, uint32_t tile, uint32_t stage,
                                                uint16_t* list_dst = nullptr, uint32_t list_count = 0) {
    const uint32_t bytes = tile_bytes_rounded(W, tile);
    const size_t origin = (size_t)tile * kTile;
    // the halo is the end of the previous tile (an L2 hit: this CTA or its neighbour just read it)
    const uint32_t halo = (kWithHalo && origin >= (size_t)kHalo) ? (uint32_t)kHalo : 0u;
    const uint32_t list_bytes = (list_count * 2u + 15u) & ~15u;
    mbar_expect_tx(&S.full_bar[stage], bytes + halo + list_bytes);
    tma_load_1d(S.data[stage] + (kHalo - halo), W.base + origin - halo, bytes + halo, &S.full_bar[stage]);
    if (list_bytes != 0u)
        tma_load_1d(list_dst, W.nl_list + (size_t)(tile - W.first_tile) * kNlCap, list_bytes, &S.full_bar[stage]);
}

// ------------------------------------------------------------------------------------------------
// front end: bitmaps + block scan + ordered newline list
// ------------------------------------------------------------------------------------------------

// 16 flag bits (byte order) of one 16-byte chunk from four lane-flag words
__device__ __forceinline__ uint32_t mask16(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    uint32_t m = bsq_gather_top(f3) >> 28;
    m = __funnelshift_l(bsq_gather_top(f2), m, 4);
    m = __funnelshift_l(bsq_gather_top(f1), m, 4);
    m = __funnelshift_l(bsq_gather_top(f0), m, 4);
    return m;
}

template <bool kHi, bool kBad, bool kEdge, typename SM>
__device__ __forceinline__ void build_bitmaps_t(SM& S, const TileCursor& c, uint32_t addlo, uint32_t addup) {
    const uint8_t* tile = S.data[c.stage] + kHalo;
    const uint32_t tid = threadIdx.x;
    const uint32_t vlo = c.lo - c.origin, vhi = c.hi - c.origin;  // valid offsets in the tile
#pragma unroll
    for (int j = 0; j < kChunksPerThread; ++j) {
        const uint32_t chunk = j * kThreads + tid;
        const uint32_t off = chunk * 16u;
        uint32_t m_nl = 0, m_hi = 0, m_bad = 0;
        if (!kEdge || (off < vhi && off + 16u > vlo)) {
            const uint4 v = *reinterpret_cast<const uint4*>(tile + off);
            m_nl = mask16(bsq_nl_flags(v.x), bsq_nl_flags(v.y), bsq_nl_flags(v.z), bsq_nl_flags(v.w));
            if (kHi) m_hi = mask16(bsq_hi_flags(v.x), bsq_hi_flags(v.y), bsq_hi_flags(v.z), bsq_hi_flags(v.w));
            if (kBad)
                m_bad = mask16(bsq_badq_flags(v.x, addlo, addup), bsq_badq_flags(v.y, addlo, addup),
                               bsq_badq_flags(v.z, addlo, addup), bsq_badq_flags(v.w, addlo, addup));
            if (kEdge && (off < vlo || off + 16u > vhi)) {
                uint32_t keep = 0xFFFFu;
                if (off < vlo) keep &= 0xFFFFu << (vlo - off);
                if (off + 16u > vhi) keep &= 0xFFFFu >> (off + 16u - vhi);
                m_nl &= keep; m_hi &= keep; m_bad &= keep;
            }
        }
        // two adjacent lanes hold the two halves of one 32-bit bitmap word
        const uint32_t x = m_nl | (m_hi << 16);
        const uint32_t xo = __shfl_down_sync(0xFFFFFFFFu, x, 1);
        uint32_t bo = 0;
        if (kBad) bo = __shfl_down_sync(0xFFFFFFFFu, m_bad, 1);
        if ((tid & 1u) == 0u) {
            const uint32_t w = chunk >> 1;
            S.bm_nl[w] = (x & 0xFFFFu) | (xo << 16);
            if (kHi) S.bm_hi_p()[w] = (x >> 16) | (xo & 0xFFFF0000u);
            if (kBad) S.bm_bad_p()[w] = m_bad | (bo << 16);
        }
    }
}
// (only the first and the last tile of a window have bytes outside [begin, end): the others skip the edge tests)
template <bool kHi, bool kBad, typename SM>
__device__ __forceinline__ void build_bitmaps(SM& S, const TileCursor& c, uint32_t addlo, uint32_t addup) {
    if (c.lo == c.origin && c.hi == c.origin + (uint32_t)kTile) build_bitmaps_t<kHi, kBad, false>(S, c, addlo, addup);
    else build_bitmaps_t<kHi, kBad, true>(S, c, addlo, addup);
}

// k_summarize with validation configured: the newline bitmap as above, plus a SCREEN of the tile for the two
// validators -- does any byte have bit 7 set / does any non-newline byte lie outside [lower, upper]?  Returns
// bit 0 / bit 1 for this thread's chunks.  No gathers and no bitmap stores: a clean tile (the normal case)
// lets k_resolve skip the validation bitmaps and the per-byte walk; a flagged tile is examined there exactly.
template <bool kHi, bool kBad, bool kEdge, typename SM>
__device__ __forceinline__ uint32_t build_nl_bitmap_and_screen_t(SM& S, const TileCursor& c, uint32_t addlo, uint32_t addup) {
    const uint8_t* tile = S.data[c.stage] + kHalo;
    const uint32_t tid = threadIdx.x;
    const uint32_t vlo = c.lo - c.origin, vhi = c.hi - c.origin;  // valid offsets in the tile
    uint32_t any_hi = 0, any_bad = 0;
#pragma unroll
    for (int j = 0; j < kChunksPerThread; ++j) {
        const uint32_t chunk = j * kThreads + tid;
        const uint32_t off = chunk * 16u;
        uint32_t m_nl = 0;
        if (!kEdge || (off < vhi && off + 16u > vlo)) {
            const uint4 v = *reinterpret_cast<const uint4*>(tile + off);
            const uint32_t f0 = bsq_nl_flags(v.x), f1 = bsq_nl_flags(v.y), f2 = bsq_nl_flags(v.z), f3 = bsq_nl_flags(v.w);
            m_nl = mask16(f0, f1, f2, f3);
            if (kHi) any_hi |= (v.x | v.y | v.z | v.w) & 0x80808080u;
            if (kBad)
                any_bad |= (bsq_badq_flags(v.x, addlo, addup) & ~f0) | (bsq_badq_flags(v.y, addlo, addup) & ~f1) |
                           (bsq_badq_flags(v.z, addlo, addup) & ~f2) | (bsq_badq_flags(v.w, addlo, addup) & ~f3);
            if (kEdge && (off < vlo || off + 16u > vhi)) {
                uint32_t keep = 0xFFFFu;
                if (off < vlo) keep &= 0xFFFFu << (vlo - off);
                if (off + 16u > vhi) keep &= 0xFFFFu >> (off + 16u - vhi);
                m_nl &= keep;
            }
        }
        const uint32_t xo = __shfl_down_sync(0xFFFFFFFFu, m_nl, 1);
        if ((tid & 1u) == 0u) S.bm_nl[chunk >> 1] = m_nl | (xo << 16);
    }
    // (bytes outside the window in the two edge tiles are not masked here: the caller flags those tiles)
    return (any_hi != 0u ? 1u : 0u) | (any_bad != 0u ? 2u : 0u);
}
template <bool kHi, bool kBad, typename SM>
__device__ __forceinline__ uint32_t build_nl_bitmap_and_screen(SM& S, const TileCursor& c, uint32_t addlo, uint32_t addup) {
    if (c.lo == c.origin && c.hi == c.origin + (uint32_t)kTile) return build_nl_bitmap_and_screen_t<kHi, kBad, false>(S, c, addlo, addup);
    return build_nl_bitmap_and_screen_t<kHi, kBad, true>(S, c, addlo, addup);
}

// Exclusive prefix of `v` over the block (in thread order) and the block total.  One barrier:
// the warp totals are double buffered (`par` alternates between consecutive scans of a CTA).
template <typename SM>
__device__ __forceinline__ uint32_t block_exclusive_scan(SM& S, uint32_t v, uint32_t& total, uint32_t& par) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    par ^= 1u;
    if (lane == 31u) S.warp_tot[par][warp][0] = inc;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t t = S.warp_tot[par][w][0];
        if ((uint32_t)w < warp) before += t;
        tot += t;
    }
    total = tot;
    return before + inc - v;
}

// Same for three values at once (the id / seq / qual streams).
__device__ __forceinline__ void block_exclusive_scan3(TileSmem& S, uint32_t v0, uint32_t v1, uint32_t v2,
                                                      uint32_t& e0, uint32_t& e1, uint32_t& e2,
                                                      uint32_t& t0, uint32_t& t1, uint32_t& t2, uint32_t& par) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t i0 = v0, i1 = v1, i2 = v2;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, i0, d);
        const uint32_t b = __shfl_up_sync(0xFFFFFFFFu, i1, d);
        const uint32_t c = __shfl_up_sync(0xFFFFFFFFu, i2, d);
        if (lane >= (uint32_t)d) { i0 += a; i1 += b; i2 += c; }
    }
    par ^= 1u;
    if (lane == 31u) { S.warp_tot[par][warp][0] = i0; S.warp_tot[par][warp][1] = i1; S.warp_tot[par][warp][2] = i2; }
    __syncthreads();
    uint32_t b0 = 0, b1 = 0, b2 = 0, s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t x0 = S.warp_tot[par][w][0], x1 = S.warp_tot[par][w][1], x2 = S.warp_tot[par][w][2];
        if ((uint32_t)w < warp) { b0 += x0; b1 += x1; b2 += x2; }
        s0 += x0; s1 += x1; s2 += x2;
    }
    e0 = b0 + i0 - v0; e1 = b1 + i1 - v1; e2 = b2 + i2 - v2;
    t0 = s0; t1 = s1; t2 = s2;
}

template <typename SM>
__device__ __forceinline__ NlWords load_nl_words(const SM& S, uint32_t tid) {
    NlWords r;
    if (kWordsPerThread == 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(&S.bm_nl[tid * 4]);
        r.w[0] = v.x; r.w[1] = v.y; r.w[kWordsPerThread - 2] = v.z; r.w[kWordsPerThread - 1] = v.w;
    } else {
        const uint2 v = *reinterpret_cast<const uint2*>(&S.bm_nl[tid * 2]);
        r.w[0] = v.x; r.w[1] = v.y;
    }
    return r;
}
__device__ __forceinline__ uint32_t popc_words(const NlWords& x) {
    uint32_t n = 0;
#pragma unroll
    for (int i = 0; i < kWordsPerThread; ++i) n += __popc(x.w[i]);
    return n;
}

// Visits the newlines of this thread's bytes in order: f(rank, position).  Most 32-byte words
// hold none, one or two newlines ("\n+\n"): the first and the last set bit are found without a
// loop; only words with three or more take the loop.  (k_resolve: the body is one shared store.)
template <typename F>
__device__ __forceinline__ void for_each_newline(const NlWords& words, uint32_t pos0, uint32_t excl, F&& f) {
    uint32_t r = excl;
#pragma unroll
    for (int i = 0; i < kWordsPerThread; ++i) {
        uint32_t m = words.w[i];
        if (m) {
            const uint32_t lsb = m & (0u - m);
            f(r++, pos0 + (31u - __clz(lsb)));
            m ^= lsb;
            if (m) {
                const uint32_t top = 31u - __clz(m);
                m ^= 1u << top;
                while (m) {                       // rare: >= 3 newlines within 32 bytes
                    const uint32_t l2 = m & (0u - m);
                    f(r++, pos0 + (31u - __clz(l2)));
                    m ^= l2;
                }
                f(r++, pos0 + top);
            }
        }
        pos0 += 32u;
    }
}
// The same visit as one plain loop per word (k_summarize: a heavier body, instantiated once).
template <typename F>
__device__ __forceinline__ void for_each_newline_loop(const NlWords& words, uint32_t pos0, uint32_t excl, F&& f) {
    uint32_t r = excl;
#pragma unroll
    for (int i = 0; i < kWordsPerThread; ++i) {
        uint32_t m = words.w[i];
        while (m) {
            f(r++, pos0 + (uint32_t)__ffs((int)m) - 1u);
            m &= m - 1u;
        }
        pos0 += 32u;
    }
}

// Writes the positions of the local newlines with rank in [pass_base, pass_base + kNlCap) to
// nlx[kHead + rank - pass_base].  `excl` = rank of the first newline of this thread's 128 bytes.
// kChecked = false when the whole tile fits one pass (the common case): no bound check per entry.
template <bool kChecked>
__device__ __forceinline__ void fill_newline_list(TileSmem& S, const TileCursor& c, const NlWords& words,
                                                  uint32_t excl, uint32_t pass_base) {
    uint32_t* const list = &S.nlx[kHead];
    for_each_newline(words, c.origin + threadIdx.x * kBytesPerThread, excl - pass_base, [&](uint32_t rel, uint32_t p) {
        if (!kChecked || rel < (uint32_t)kNlCap) list[rel] = p;
    });
}

// After a pass of n entries: the kHead most recent newline positions move to the front.
__device__ __forceinline__ void rotate_head(TileSmem& S, uint32_t n) {
    // caller guarantees a barrier before (all readers done) and after
    if (threadIdx.x == 0) {
        uint32_t t[kHead];
#pragma unroll
        for (int i = 0; i < kHead; ++i) t[i] = S.nlx[n + i];
#pragma unroll
        for (int i = 0; i < kHead; ++i) S.nlx[i] = t[i];
    }
}

// byte of the window at offset pos: shared memory when the current tile holds it
__device__ __forceinline__ uint32_t byte_at(const TileSmem& S, const TileCursor& c, const WinParams& W,
                                            uint32_t pos) {
    const uint32_t rel = pos - c.origin + (uint32_t)kHalo;   // offset in data[stage]; wraps for pos far before
    if (rel < (uint32_t)(kHalo + kTile) && pos + (uint32_t)kHalo >= c.origin) return S.data[c.stage][rel];
    return __ldg(W.base + pos);
}

// ------------------------------------------------------------------------------------------------
// k_summarize
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void run_tiles(const WinParams& W, uint32_t run, uint32_t& ta, uint32_t& tb) {
    ta = W.first_tile + run * W.tiles_per_run;
    tb = ta + W.tiles_per_run;
    if (tb > W.n_tiles) tb = W.n_tiles;
    if (ta > tb) ta = tb;
}

__device__ __forceinline__ TileCursor make_cursor(const WinParams& W, uint32_t tile, uint32_t stage) {
    TileCursor c;
    c.tile = tile;
    c.origin = tile * (uint32_t)kTile;
    c.lo = c.origin < W.begin ? W.begin : c.origin;
    const uint32_t e = c.origin + (uint32_t)kTile;
    c.hi = (e > W.end || e < c.origin) ? W.end : e;
    c.stage = stage;
    return c;
}

// kSums = false (views-only passes): only the newline count and the first/last positions of the
// run are needed; the per-class position sums that give the SoA destinations are skipped.
// Shared memory: SumSmem.
constexpr uint32_t kNlCountMask = 0x0FFFFFFFu;   // nl_count word: newlines | validation screen << 30
template <bool kSums, bool kHi, bool kBad>
__global__ void __launch_bounds__(kThreads, SumShape<kSums>::kCtas) k_summarize(const WinParams W, BsqSummary* __restrict__ run_sum,
                                                                        uint32_t lower, uint32_t upper) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int kSumStages = SumShape<kSums>::kStagesOf;
    using SumSmem = SumSmemT<kSumStages>;
    SumSmem& S = *reinterpret_cast<SumSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    uint32_t ta, tb;
    run_tiles(W, blockIdx.x, ta, tb);

    if (tid == 0) {
        for (int s = 0; s < kSumStages; ++s) mbar_init(&S.full_bar[s], 1);
        mbar_fence_init();
    }
    if (tid < 8) S.k1_head[tid] = 0;
    if (tid == 0) S.dirty = 0;
    __syncthreads();
    if (tid == 0)
        for (uint32_t s = 0; s < (uint32_t)kSumStages && ta + s < tb; ++s) issue_tile_load<false>(S, W, ta + s, s);
    const uint32_t addlo = (128u - lower) * 0x01010101u, addup = (127u - upper) * 0x01010101u;

    uint32_t run_count = 0;           // newlines of the run so far (uniform)
    uint32_t acc[4] = {0, 0, 0, 0};   // position sums by (index in run) mod 4, this thread's share
    uint32_t par = 0;

    for (uint32_t t = ta; t < tb; ++t) {
        const uint32_t it = t - ta;
        const TileCursor c = make_cursor(W, t, it % kSumStages);
        mbar_wait(&S.full_bar[c.stage], (it / kSumStages) & 1u);
        if (kSumStages == 1 && tid == 0 && t + 1u < tb) prefetch_l2(W.base + (size_t)(t + 1u) * kTile, tile_bytes_rounded(W, t + 1u));
        if (kHi || kBad) {
            uint32_t d = build_nl_bitmap_and_screen<kHi, kBad>(S, c, addlo, addup);
            d = __reduce_or_sync(0xFFFFFFFFu, d);
            if (d != 0u && (tid & 31u) == 0u) atomicOr(&S.dirty, d);
        } else {
            build_bitmaps<false, false>(S, c, 0, 0);
        }
        __syncthreads();
        const NlWords words = load_nl_words(S, tid);
        const uint32_t cnt = popc_words(words);
        uint32_t total;
        const uint32_t excl = block_exclusive_scan(S, cnt, total, par);
        if (total <= (uint32_t)kNlCap) {
            // the ordered list of the tile's newlines (tile-relative, 16 bit): k_resolve picks it up instead of
            // rebuilding it, and the run's position sums and edge positions are read off it here
            uint16_t* const list = S.list();
            if (cnt != 0u)
                for_each_newline(words, tid * kBytesPerThread, excl, [&](uint32_t r, uint32_t p) { list[r] = (uint16_t)p; });
            __syncthreads();
            if (kSums) {
                // index in the run = run_count + j: thread tid adds entries 2 tid and 2 tid + 1 (+ 2 kThreads, ...),
                // i.e. two fixed classes; the entries past `total` of the last pair are not counted
                const uint32_t* l2 = reinterpret_cast<const uint32_t*>(list);
                uint32_t a0 = 0, a1 = 0;
                for (uint32_t j = 2u * tid; j < total; j += 2u * kThreads) {
                    const uint32_t w = l2[j >> 1];
                    a0 += c.origin + (w & 0xFFFFu);
                    a1 += j + 1u < total ? c.origin + (w >> 16) : 0u;
                }
                const uint32_t cls = (run_count + 2u * tid) & 3u;    // class of the even entry; the odd one is cls + 1
                acc[0] += cls == 0u ? a0 : (cls == 3u ? a1 : 0u); acc[1] += cls == 1u ? a0 : (cls == 0u ? a1 : 0u);
                acc[2] += cls == 2u ? a0 : (cls == 1u ? a1 : 0u); acc[3] += cls == 3u ? a0 : (cls == 2u ? a1 : 0u);
            }
            if (W.nl_list != nullptr) {                // coalesced copy-out, two positions per store
                uint32_t* g = reinterpret_cast<uint32_t*>(W.nl_list + (size_t)(t - W.first_tile) * kNlCap);
                const uint32_t* l = reinterpret_cast<const uint32_t*>(list);
                for (uint32_t j = tid; j < (total + 1u) >> 1; j += kThreads) g[j] = l[j];
            }
            if (tid == 0 && total != 0u) {
                for (uint32_t r = 0; r < total && run_count + r < 4u; ++r)           // first[] of the run
                    S.k1_head[4u + run_count + r] = c.origin + (uint32_t)list[r];
                // merge the tile's last (up to four) newlines into the run's last four (oldest first)
                const uint32_t k = total < 4u ? total : 4u;
                uint32_t h[4];
                for (uint32_t i = 0; i < 4u; ++i)
                    h[i] = i + k < 4u ? S.k1_head[i + k] : c.origin + (uint32_t)list[total - 4u + i];
                for (uint32_t i = 0; i < 4u; ++i) S.k1_head[i] = h[i];
            }
        } else {
            // more newlines than the list holds: visit them from the bitmap
            const bool edge = run_count + excl < 4u || excl + cnt + 4u > total;   // among the first / last four
            if (cnt != 0u && (kSums || edge)) {
                for_each_newline_loop(words, c.origin + tid * kBytesPerThread, excl, [&](uint32_t r, uint32_t p) {
                    if (kSums) {
                        const uint32_t cls = (run_count + r) & 3u;
                        acc[0] += cls == 0u ? p : 0u;
                        acc[1] += cls == 1u ? p : 0u;
                        acc[2] += cls == 2u ? p : 0u;
                        acc[3] += cls == 3u ? p : 0u;
                    }
                    if (edge) {
                        if (run_count + r < 4u) S.k1_head[4u + run_count + r] = p;     // first[] of the run
                        if (r + 4u >= total) S.carry[r + 4u - total] = p;             // last four of the tile
                    }
                });
            }
            __syncthreads();
            if (tid == 0 && total != 0u) {
                const uint32_t k = total < 4u ? total : 4u;
                uint32_t h[4];
                for (uint32_t i = 0; i < 4u; ++i) h[i] = i + k < 4u ? S.k1_head[i + k] : S.carry[i];
                for (uint32_t i = 0; i < 4u; ++i) S.k1_head[i] = h[i];
            }
        }
        if (tid == 0) {
            // an edge tile of the window holds foreign bytes the screen did not mask: examine it in k_resolve
            const bool edge_tile = c.lo != c.origin || c.hi != c.origin + (uint32_t)kTile;
            const uint32_t dirty = (kHi || kBad) ? (edge_tile ? 3u : S.dirty) : 0u;
            if (W.nl_count != nullptr) W.nl_count[t - W.first_tile] = total | (dirty << 30);
            S.dirty = 0;                               // (read above; the next tile ORs into it after the barrier below)
        }
        run_count += total;
        __syncthreads();  // every thread is done with data[stage] and the head is updated
        if (tid == 0 && t + kSumStages < tb) issue_tile_load<false>(S, W, t + kSumStages, c.stage);
    }

    // block reduction of acc[4]
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += __shfl_xor_sync(0xFFFFFFFFu, acc[k], d);
    }
    __syncthreads();
    if ((tid & 31u) == 0u)
        for (int k = 0; k < 4; ++k) S.k1_red[(tid >> 5) * 4 + k] = acc[k];
    __syncthreads();
    if (tid == 0) {
        BsqSummary s = bsq_summary_identity();
        s.count = run_count;
        for (int w = 0; w < kWarps; ++w)
            for (int k = 0; k < 4; ++k) s.P[k] += S.k1_red[w * 4 + k];
        for (int i = 0; i < 4; ++i) s.last[i] = S.k1_head[3 - i];
        for (uint32_t i = 0; i < 4u; ++i) s.first[i] = i < run_count ? S.k1_head[4u + i] : 0u;
        run_sum[blockIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// k_scan_runs: one CTA; n_runs is a few hundred
// ------------------------------------------------------------------------------------------------

constexpr int kMaxRuns = 2048;
constexpr int kScanThreads = 1024;

__device__ __forceinline__ LbState lb_from_summary(const BsqSummary& s) {
    LbState v;
    v.count = s.count;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v.last[i] = s.last[i]; v.P[i] = s.P[i]; }
    return v;
}
__device__ __forceinline__ LbState lb_shfl_up(const LbState& v, uint32_t d) {
    LbState o;
    o.count = __shfl_up_sync(0xFFFFFFFFu, v.count, d);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o.last[i] = __shfl_up_sync(0xFFFFFFFFu, v.last[i], d);
        o.P[i] = __shfl_up_sync(0xFFFFFFFFu, v.P[i], d);
    }
    return o;
}
// inclusive scan over the lanes of a warp (lane order = stream order)
__device__ __forceinline__ LbState lb_warp_inclusive(LbState v) {
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (uint32_t d = 1; d < 32u; d <<= 1) {
        const LbState o = lb_shfl_up(v, d);
        if (lane >= d) v = lb_combine(o, v);
    }
    return v;
}

// One CTA of 1024 threads: every thread folds its (<= 2) runs, a shuffle scan per warp, a shuffle scan of
// the 32 warp totals, then every thread emits the prefixes of its runs.  Works on the nine words of a
// summary that cross runs (LbState); the window-init state is always the leftmost operand.
__global__ void __launch_bounds__(kScanThreads, 1) k_scan_runs(const BsqSummary* __restrict__ run_sum, uint32_t n_runs,
                                                               uint32_t begin, BsqPrefix* __restrict__ run_pre,
                                                               ScanOut* __restrict__ out) {
    __shared__ LbState s_warp[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t per = (n_runs + kScanThreads - 1u) / kScanThreads;
    const uint32_t r0 = tid * per < n_runs ? tid * per : n_runs, r1 = r0 + per < n_runs ? r0 + per : n_runs;
    LbState g = lb_identity();
    for (uint32_t r = r0; r < r1; ++r) g = lb_combine(g, lb_from_summary(run_sum[r]));
    const LbState inc = lb_warp_inclusive(g);
    if (lane == 31u) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const LbState w = lb_warp_inclusive(s_warp[lane]);
        LbState ex = lb_shfl_up(w, 1);
        if (lane == 0) ex = lb_identity();
        s_warp[lane] = ex;                         // runs of the warps before this one
        if (lane == 31u) {
            // totals of the window
            LbState init = lb_identity();
            init.last[0] = begin - 1u;             // bsq_summary_window_init
            const BsqSummary E = lb_to_summary(lb_combine(init, w));
            out->totals = bsq_totals_from(E, begin);
            out->end_state = E;
            // the region without the window init (shard stitching); first[] from the leading runs
            BsqSummary R = bsq_summary_identity();
            for (uint32_t r = 0; r < n_runs && R.count < 4u; ++r) R = bsq_combine(R, run_sum[r]);
            BsqSummary reg = lb_to_summary(w);
            for (int i = 0; i < 4; ++i) reg.first[i] = R.first[i];
            out->region = reg;
        }
    }
    __syncthreads();
    LbState before = lb_shfl_up(inc, 1);           // runs of the lower lanes of this warp
    if (lane == 0) before = lb_identity();
    LbState init = lb_identity();
    init.last[0] = begin - 1u;
    LbState E = lb_combine(lb_combine(init, s_warp[warp]), before);
    for (uint32_t r = r0; r < r1; ++r) {
        run_pre[r] = bsq_prefix_from(lb_to_summary(E), begin);
        E = lb_combine(E, lb_from_summary(run_sum[r]));
    }
}

// ------------------------------------------------------------------------------------------------
// k_resolve
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void report(const ResolveParams& P, uint32_t k, uint32_t code) {
    const unsigned long long key = ((unsigned long long)(P.first_record + P.rec_base + (int64_t)k) << 8) | code;
    atomicMin(P.err, key);
}

// 16 source bytes starting at window offset pos (any alignment; pos may wrap below 0 when a line
// begins inside the destination vector -- those leading bytes are never used).  The current tile is
// read from shared memory (two aligned 16-byte loads + funnel shifts); bytes of a line that began
// in an earlier tile come from global memory (L2), word by word, guarded to the window.
__device__ __forceinline__ uint4 load16(const TileSmem& S, const TileCursor& c, const WinParams& W, uint32_t pos) {
    const uint32_t rel = pos - c.origin + (uint32_t)kHalo;   // offset in data[stage] = [halo | tile | pad]
    const uint32_t sh = (pos & 3u) * 8u;
    const bool in_tile = rel < (uint32_t)(kHalo + kTile) && (c.origin >= (uint32_t)kHalo || rel >= (uint32_t)kHalo - 16u);
    // five consecutive words: bank conflicts, but no selects.  The read is unconditional (address
    // clamped into the buffer) so that the common path has no branch.
    const uint32_t* t32 = reinterpret_cast<const uint32_t*>(S.data[c.stage]) + (in_tile ? (rel >> 2) : 0u);
    uint32_t w0 = t32[0], w1 = t32[1], w2 = t32[2], w3 = t32[3], w4 = t32[4];
    if (!in_tile) {   // the line began more than kHalo before the tile: global memory (L2), guarded to the window
        const uint32_t* g = reinterpret_cast<const uint32_t*>(W.base);
        const uint32_t wi = pos >> 2;                       // garbage when pos wrapped: guarded
        const uint32_t wend = (W.end + 3u) >> 2;            // words that hold window bytes
        const bool neg = (int32_t)pos < 0;
        w0 = (!neg && wi + 0u < wend) ? __ldg(g + wi + 0u) : 0u;
        w1 = (!neg && wi + 1u < wend) ? __ldg(g + wi + 1u) : 0u;
        w2 = (!neg && wi + 2u < wend) ? __ldg(g + wi + 2u) : 0u;
        w3 = (!neg && wi + 3u < wend) ? __ldg(g + wi + 3u) : 0u;
        w4 = (!neg && wi + 4u < wend) ? __ldg(g + wi + 4u) : 0u;
    }
    uint4 r;
    r.x = __funnelshift_r(w0, w1, sh);
    r.y = __funnelshift_r(w1, w2, sh);
    r.z = __funnelshift_r(w2, w3, sh);
    r.w = __funnelshift_r(w3, w4, sh);
    return r;
}

// bytes [0, a) of the result come from acc, bytes [a, 16) from x   (0 < a <= 16)
__device__ __forceinline__ uint4 splice16(uint4 acc, uint4 x, uint32_t a) {
    // per word: the k = clamp(a - 4w, 0, 4) low bytes stay from acc; mask = 0xFFFFFFFF >> 8(4 - k),
    // with the shift clamped at 32 (funnel shift) so that k == 0 gives 0
    auto keep = [&](int w) -> uint32_t {
        int k = (int)a - 4 * w;
        k = k < 0 ? 0 : (k > 4 ? 4 : k);
        return __funnelshift_rc(0xFFFFFFFFu, 0u, 8u * (uint32_t)(4 - k));
    };
    const uint32_t m0 = keep(0), m1 = keep(1), m2 = keep(2), m3 = keep(3);
    acc.x = (acc.x & m0) | (x.x & ~m0);
    acc.y = (acc.y & m1) | (x.y & ~m1);
    acc.z = (acc.z & m2) | (x.z & ~m2);
    acc.w = (acc.w & m3) | (x.w & ~m3);
    return acc;
}

// stores bytes [a, b) of v at p[a..b)   (p 16-byte aligned; used for the two edge vectors of a range)
__device__ __forceinline__ void store_partial16(uint8_t* p, uint4 v, uint32_t a, uint32_t b) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (uint32_t i = 0; i < 4u; ++i) {
        if (a <= 4u * i && b >= 4u * i + 4u) {
            *reinterpret_cast<uint32_t*>(p + 4u * i) = w[i];
        } else {
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k)
                if (4u * i + k >= a && 4u * i + k < b) p[4u * i + k] = (uint8_t)(w[i] >> (8u * k));
        }
    }
}

struct StreamJob {
    const uint32_t* sdst; const uint32_t* ssrc;
    uint32_t n_lines, d0, d1; uint8_t* out;
};

// the 16 bytes of destination vector v of a stream (byte k <-> destination 16 v + k)
__device__ __forceinline__ uint4 assemble_vector(const TileSmem& S, const TileCursor& c, const WinParams& W,
                                                 const StreamJob& J, uint32_t v, uint32_t lo, uint32_t hi) {
    const uint32_t vs = v * 16u;
    uint32_t i;
    {                                          // line that holds the first byte of the vector: bisect
        uint32_t a = 0, b = J.n_lines;
        while (b - a > 1u) {
            const uint32_t m = (a + b) >> 1;
            if (J.sdst[m] <= lo) a = m; else b = m;
        }
        i = a;
    }
    uint32_t e = J.sdst[i + 1];
    const uint32_t a1 = e - vs < 16u ? e - vs : 16u;              // bytes of this vector before line i+1
    uint4 acc = load16(S, c, W, vs + (J.ssrc[i] - J.sdst[i]));
    acc = splice16(acc, load16(S, c, W, J.ssrc[i + 1] - a1), a1);
    if (e < hi) {
        ++i;
        e = J.sdst[i + 1];
        while (e < hi) {                       // a third (fourth, ...) line begins inside this vector
            ++i;
            const uint32_t a = e - vs;
            if (J.sdst[i + 1] != e) acc = splice16(acc, load16(S, c, W, J.ssrc[i] - a), a);
            e = J.sdst[i + 1];
        }
    }
    return acc;
}

// Copies the lines of one class stream that ended in this pass: destination range [d0, d1) of the
// stream (virtual offsets: `out` is 16-byte aligned and d includes the sub-16 shift).  Source of
// destination byte d inside line i is d + (ssrc[i] - sdst[i]), so a 16-byte destination vector is
// two unaligned 16-byte reads (its first line and the next one) spliced at the line boundary and
// one aligned 16-byte store; a third line inside the same vector (ids, very short reads) takes a
// loop.  Two vectors per thread are in flight.
__device__ __noinline__ void copy_stream(const TileSmem& S, const TileCursor& c, const WinParams& W, const StreamJob J) {
    const uint32_t v0 = J.d0 >> 4, v1 = (J.d1 + 15u) >> 4;
    for (uint32_t v = v0 + threadIdx.x; v < v1; v += 2u * kThreads) {
        const uint32_t u = v + kThreads;
        const bool two = u < v1;
        const uint32_t lo_v = v * 16u < J.d0 ? J.d0 : v * 16u, hi_v = v * 16u + 16u > J.d1 ? J.d1 : v * 16u + 16u;
        const uint32_t lo_u = u * 16u, hi_u = u * 16u + 16u > J.d1 ? J.d1 : u * 16u + 16u;
        const uint4 av = assemble_vector(S, c, W, J, v, lo_v, hi_v);
        uint4 au = make_uint4(0, 0, 0, 0);
        if (two) au = assemble_vector(S, c, W, J, u, lo_u, hi_u);
        uint8_t* pv = J.out + (size_t)v * 16u;
        if (hi_v - lo_v == 16u) *reinterpret_cast<uint4*>(pv) = av;
        else store_partial16(pv, av, lo_v - v * 16u, hi_v - v * 16u);
        if (two) {
            uint8_t* pu = J.out + (size_t)u * 16u;
            if (hi_u - lo_u == 16u) *reinterpret_cast<uint4*>(pu) = au;
            else store_partial16(pu, au, 0u, hi_u - lo_u);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// direct SoA copy: the lines that end in a tile go from [halo | tile] in shared memory straight to their
// place in the global arenas (no staging buffer).
//
// The copy is organised by DESTINATION, not by line: consecutive lanes write consecutive 16-byte
// vectors of a stream (one coalesced 512-byte store per warp), and their sources are consecutive
// 16-byte windows of one line in shared memory, read as the two aligned 16-byte vectors that hold
// them -- so neither side has a bank-conflict pattern that depends on the record stride (a per-line
// word loop puts lane i at i x stride: 16-way conflicts when the stride is 320 bytes).
//   interior   every destination vector that lies inside ONE line: one thread per vector
//   edges      a vector in which a line ends (it also takes the first bytes of the following
//              line(s)) and the partial first / last vector of the tile's range: one thread per line
//   ids        short lines, one thread per line, word stores
// ------------------------------------------------------------------------------------------------

// 16 bytes at byte offset `src` of `data` (any alignment; `data` is 16-byte aligned and has 32 readable
// bytes past the last one asked for)
__device__ __forceinline__ uint4 load16_smem(const uint8_t* __restrict__ data, uint32_t src) {
    const uint4* q = reinterpret_cast<const uint4*>(data + (src & ~15u));
    const uint4 a = q[0], b = q[1];
    uint32_t v0 = a.x, v1 = a.y, v2 = a.z, v3 = a.w, v4 = b.x, v5 = b.y;
    if (src & 8u) { v0 = v2; v1 = v3; v2 = v4; v3 = v5; v4 = b.z; v5 = b.w; }
    if (src & 4u) { v0 = v1; v1 = v2; v2 = v3; v3 = v4; v4 = v5; }
    const uint32_t sh = (src & 3u) * 8u;
    return make_uint4(__funnelshift_r(v0, v1, sh), __funnelshift_r(v1, v2, sh), __funnelshift_r(v2, v3, sh),
                      __funnelshift_r(v3, v4, sh));
}

struct DirectJob {
    const uint32_t* sdst; const uint32_t* ssrc;   // per line: destination (virtual offset) and window position
    uint32_t n_lines, ra, rb;                      // sdst[0] == ra, sdst[n_lines] == rb
    uint8_t* out;                                  // 16-byte aligned; destination byte d is out[d]
};

// interior vectors of one stream (block-wide)
__device__ __forceinline__ void copy_interior(const uint8_t* __restrict__ data, uint32_t sbias, const DirectJob& J) {
    const uint32_t v_lo = (J.ra + 15u) >> 4, v_hi = J.rb >> 4;
    if (v_hi <= v_lo) return;
    const float inv = (float)J.n_lines / (float)(J.rb - J.ra);   // lines per destination byte: first guess of the line
    for (uint32_t v = v_lo + threadIdx.x; v < v_hi; v += kThreads) {
        const uint32_t p = v * 16u;
        uint32_t i = (uint32_t)((float)(p - J.ra) * inv);
        if (i >= J.n_lines) i = J.n_lines - 1u;
        while (J.sdst[i] > p) --i;                 // sdst[0] == ra <= p
        uint32_t e = J.sdst[i + 1];
        while (e <= p) { ++i; e = J.sdst[i + 1]; } // sdst[n_lines] == rb > p; skips empty lines
        if (e >= p + 16u)                          // (a vector with a line end inside belongs to copy_edges)
            *reinterpret_cast<uint4*>(J.out + p) = load16_smem(data, p + (J.ssrc[i] - J.sdst[i]) + sbias);
    }
}

// destination vector V of a stream, assembled from line i (which holds the vector's first byte that lies in
// [ra, rb)) and as many following lines as begin inside it
__device__ __forceinline__ void copy_edge_vector(const uint8_t* __restrict__ data, uint32_t sbias, const DirectJob& J,
                                                 uint32_t i, uint32_t V) {
    const uint32_t vs = V * 16u;
    const uint32_t lo = vs < J.ra ? J.ra : vs, hi = vs + 16u > J.rb ? J.rb : vs + 16u;
    uint4 acc = load16_smem(data, vs + (J.ssrc[i] - J.sdst[i]) + sbias);
    uint32_t k = i, e = J.sdst[i + 1];
    while (e < hi) {
        ++k;
        const uint32_t a = e - vs, nx = J.sdst[k + 1];
        if (nx != e) acc = splice16(acc, load16_smem(data, J.ssrc[k] + sbias - a), a);
        e = nx;
    }
    uint8_t* p = J.out + vs;
    if (hi - lo == 16u) *reinterpret_cast<uint4*>(p) = acc;
    else store_partial16(p, acc, lo - vs, hi - vs);
}

// the edge vectors line i owns (one thread per line)
__device__ __forceinline__ void copy_edges(const uint8_t* __restrict__ data, uint32_t sbias, const DirectJob& J, uint32_t i) {
    const uint32_t s0 = J.sdst[i], s1 = J.sdst[i + 1];
    if (s1 == s0) return;                                            // an empty line owns nothing
    if (s1 & 15u) {                                                  // the line ends inside a vector: its owner is the
        const uint32_t V = s1 >> 4;                                  // line that holds the vector's first byte in range
        const uint32_t first = V * 16u < J.ra ? J.ra : V * 16u;
        if (s0 <= first) copy_edge_vector(data, sbias, J, i, V);
    }
    // the partial first vector of the range, when this line fills it to its end
    if (s0 == J.ra && (J.ra & 15u) && (s1 >> 4) != (J.ra >> 4)) copy_edge_vector(data, sbias, J, i, J.ra >> 4);
}

// a short line (ids): head bytes to the first 4-byte boundary of the destination, one funnel shift per
// word, the last 0-3 bytes
__device__ __forceinline__ void copy_small_line(const uint8_t* __restrict__ data, uint32_t src, uint8_t* __restrict__ out,
                                                uint32_t dst, uint32_t len) {
    if (len == 0u) return;   // (the source of an empty line may lie outside [halo | tile])
    uint32_t h = (0u - dst) & 3u;
    if (h > len) h = len;
    for (uint32_t i = 0; i < h; ++i) out[dst + i] = data[src + i];
    src += h; dst += h; len -= h;
    const uint32_t nw = len >> 2;
    const uint32_t sh = (src & 3u) * 8u;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(data + (src & ~3u));
    uint32_t* dw = reinterpret_cast<uint32_t*>(out + dst);
    uint32_t w0 = sw[0];
    for (uint32_t i = 0; i < nw; ++i) {
        const uint32_t w1 = sw[i + 1];
        dw[i] = __funnelshift_r(w0, w1, sh);
        w0 = w1;
    }
    const uint32_t t = len & 3u;
    src += nw * 4u; dst += nw * 4u;
    for (uint32_t k = 0; k < t; ++k) out[dst + k] = data[src + k];
}

#if BSQ_COPY_STAGED
// ------------------------------------------------------------------------------------------------
// staged SoA copy: one thread moves one line from [halo | tile] to `stage`, which is laid out like the
// destination modulo 16: a byte loop up to the first 4-byte boundary of the destination, then one funnel
// shift per word (each source word is read once), then the last 0-3 bytes.  The staged range leaves
// through TMA bulk stores, so the global writes cost no instructions and are full 16-byte vectors.
//
// Bank conflicts: lane i of a warp reads word (line start of lane i) + k at step k, so lines that start
// in the same bank collide at every step -- 16 of 32 lanes when the record stride is 320 bytes.  The
// lanes of a warp whose lines start in the same bank therefore begin at different words: lane with
// rank r among them copies words [r, nw) first and then [0, r) downwards, which keeps the carried
// word and costs r_max extra steps (7 at a 320-byte stride with the sequence and quality lines
// interleaved over the lanes).  Warps whose lanes spread over enough banks skip all of this.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_line(const uint8_t* __restrict__ data, uint8_t* __restrict__ stage,
                                           uint32_t src, uint32_t dst, uint32_t len, uint32_t lanes) {
    uint32_t h = (0u - dst) & 3u;
    if (h > len) h = len;
    const uint32_t src4 = src + h;
    // rank of this lane among the lanes of the warp whose word loop starts in the same bank
    // how many banks do the lanes' first words occupy?  Few banks for many lanes = lanes in step on the same
    // bank: those get distinct start words (rank among the lanes of their bank).  The common case (at most
    // ~2 lanes per bank) takes no rotation and pays one warp reduction.
    uint32_t rot = 0;
#if BSQ_ROT == 2
    rot = lanes;           // (the caller's arithmetic rank: see the call site)
    (void)src4;
#elif BSQ_ROT
    const uint32_t bank = (src4 >> 2) & 31u;
    const uint32_t occupied = __reduce_or_sync(lanes, 1u << bank);
    if (3u * __popc(occupied) < __popc(lanes)) {
        const uint32_t same = __match_any_sync(lanes, bank);
        rot = __popc(same & ((1u << (threadIdx.x & 31u)) - 1u));
    }
#endif
    if (len == 0u) return;   // (the source of an empty line may lie outside [halo | tile])
    for (uint32_t i = 0; i < h; ++i) stage[dst + i] = data[src + i];
    src += h; dst += h; len -= h;
    const uint32_t nw = len >> 2;
    const uint32_t sh = (src & 3u) * 8u;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(data + (src & ~3u));
    uint32_t* dw = reinterpret_cast<uint32_t*>(stage + dst);
    if (rot >= nw) rot = 0u;
    const uint32_t wr = sw[rot];
    uint32_t w0 = wr;
    uint32_t i = rot;
    for (; i + 4u <= nw; i += 4u) {
        const uint32_t w1 = sw[i + 1], w2 = sw[i + 2], w3 = sw[i + 3], w4 = sw[i + 4];
        dw[i] = __funnelshift_r(w0, w1, sh);
        dw[i + 1] = __funnelshift_r(w1, w2, sh);
        dw[i + 2] = __funnelshift_r(w2, w3, sh);
        dw[i + 3] = __funnelshift_r(w3, w4, sh);
        w0 = w4;
    }
    for (; i < nw; ++i) {
        const uint32_t w1 = sw[i + 1];
        dw[i] = __funnelshift_r(w0, w1, sh);
        w0 = w1;
    }
    if (rot != 0u) {                                    // words [0, rot): the loads first (independent), then the stores
        uint32_t hw[9];
#pragma unroll
        for (uint32_t k = 0; k < 9u; ++k) hw[k] = k <= rot ? sw[k] : 0u;
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k)
            if (k < rot) dw[k] = __funnelshift_r(hw[k], hw[k + 1], sh);
        uint32_t w1 = hw[8];
        for (uint32_t k = 8u; k < rot; ++k) {           // (more than eight lanes on one bank)
            const uint32_t w = sw[k + 1];
            dw[k] = __funnelshift_r(w1, w, sh);
            w1 = w;
        }
    }
    const uint32_t t = len & 3u;
    src += nw * 4u; dst += nw * 4u;
    for (uint32_t k = 0; k < t; ++k) stage[dst + k] = data[src + k];
}

// One stream's staged range -> global: [ra, rb) are virtual destination offsets (out + offset, out
// 16-byte aligned), stage[so + (d - (ra & ~15))] holds destination byte d.  The 16-byte aligned
// interior leaves as one TMA bulk store (thread `issuer`); the <= 15 bytes on either side, which
// share their vector with a neighbouring tile, are byte stores by lanes 0..31 of warp `wsel`.
__device__ __forceinline__ void flush_stream(const uint8_t* stage, uint32_t so, uint32_t ra, uint32_t rb, uint8_t* out,
                                             bool issuer, bool edge_warp) {
    if (rb <= ra) return;
    const uint32_t A = ra & ~15u;
    const uint32_t i0 = (ra + 15u) & ~15u, i1 = rb & ~15u;
    if (issuer && i1 > i0) tma_store_1d(out + i0, stage + so + (i0 - A), i1 - i0);
    if (edge_warp) {
        const uint32_t j = threadIdx.x & 31u;
        const uint32_t head_end = i0 < rb ? i0 : rb;                 // head = [ra, head_end)
        const uint32_t tail_beg = i1 > head_end ? i1 : head_end;     // tail = [tail_beg, rb)
        const uint32_t d = j < 16u ? ra + j : tail_beg + (j - 16u);
        const bool on = j < 16u ? d < head_end : d < rb;
        if (on) out[d] = stage[so + (d - A)];
    }
}
#endif  // BSQ_COPY_STAGED

// The second pass: run prefixes from k_summarize + k_scan_runs, one CTA per run.
template <bool kAscii, bool kQual, bool kOffsets, bool kPack>
__global__ void __launch_bounds__(kThreads, kPack ? kResolveCtas : kViewCtas) k_resolve(const WinParams W, const ResolveParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TileSmem& S = *reinterpret_cast<TileSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    uint32_t ta = 0, tb = 0;
    run_tiles(W, blockIdx.x, ta, tb);
    const BsqPrefix pre = P.run_pre[blockIdx.x];

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&S.full_bar[s], 1);
        mbar_fence_init();
        S.nlx[0] = pre.prev3; S.nlx[1] = pre.prev[2]; S.nlx[2] = pre.prev[1]; S.nlx[3] = pre.prev[0];
    }
    __syncthreads();
    // use_list: the ordered newline list of every tile comes from k_summarize (same TMA transaction as the tile),
    // so the bitmap / scan / list front end runs only where the bitmaps are needed anyway (a tile that
    // k_summarize's validation screen flagged) and for tiles with more newlines than the list holds
    const bool use_list = W.nl_list != nullptr;
    auto list_count_of = [&](uint32_t tile) -> uint32_t { return W.nl_count[tile - W.first_tile]; };   // raw word
    // newlines of a tile whose list is usable by this instantiation, else 0 (raw = count | screen << 30)
    auto list_ok = [&](uint32_t raw) -> bool {
        const uint32_t n = raw & kNlCountMask, dirty = raw >> 30;
        return n <= (uint32_t)kNlCap && !((kAscii && (dirty & 1u)) || (kQual && (dirty & 2u)));
    };
    auto listed_count = [&](uint32_t raw) -> uint32_t { return list_ok(raw) ? (raw & kNlCountMask) : 0u; };
    // thread 0 runs the tile ring: tile t + kStages is requested when tile t is done; the newline counts travel
    // one tile further ahead (a global load whose latency nobody waits for)
    uint32_t n_ahead = 0;                              // (thread 0) raw count word of tile t + kStages
    if (tid == 0) {
        for (uint32_t s = 0; s < (uint32_t)kStages && ta + s < tb; ++s) {
            uint32_t raw = 0;
            if (use_list) { raw = list_count_of(ta + s); S.tile_total[s] = raw; }
            issue_tile_load<true>(S, W, ta + s, s, S.nl16[s], listed_count(raw));
        }
        if (use_list && ta + (uint32_t)kStages < tb) n_ahead = list_count_of(ta + (uint32_t)kStages);
    }
    if (kOffsets && tid == 0 && blockIdx.x == 0) P.line_ends[0] = W.begin - 1u;

    // (q5_width: k_summarize's screen counts UPPER itself as suspicious; whether it is an error is decided per record)
    const uint32_t up_bm = P.upper - (kQual && P.q5_width != 0u ? 1u : 0u);
    const uint32_t addlo = (128u - P.lower) * 0x01010101u, addup = (127u - up_bm) * 0x01010101u;
    uint32_t bases_acc = 0;                                              // this thread's share of sum(seq_len)
    uint32_t rank = pre.rank;                                            // rank of the tile's first newline
    uint32_t cum_id = pre.cum_id, cum_seq = pre.cum_seq, cum_qual = pre.cum_qual;  // stream destinations
    uint32_t par = 0;
    // virtual destination offsets: out pointers rounded down to 16 bytes, offsets shifted up
    const uint32_t sh_id = (uint32_t)(P.id_base64 & 15), sh_seq = (uint32_t)(P.seq_base64 & 15),
                   sh_qual = (uint32_t)(P.qual_base64 & 15);
    uint8_t* const out_id = kPack ? P.id_out + (P.id_base64 - sh_id) : nullptr;
    uint8_t* const out_seq = kPack ? P.seq_out + (P.seq_base64 - sh_seq) : nullptr;
    uint8_t* const out_qual = kPack ? P.qual_out + (P.qual_base64 - sh_qual) : nullptr;
    const uint32_t bsz = (uint32_t)P.batch_size;
    // window-local index of the next record that closes a batch (uniform; advanced as the run proceeds)
    uint32_t edge_k = (pre.rank >> 2) + (bsz - 1u - (P.rec_mod + (pre.rank >> 2)) % bsz);
    const uint32_t n_complete = P.n_complete;
    constexpr uint32_t kRing = (uint32_t)kStages + 1u;  // tile_total ring: written one tile before it is read

    for (uint32_t t = ta, it = 0; t < tb; ++t, ++it) {
        const TileCursor c = make_cursor(W, t, it % kStages);
        mbar_wait(&S.full_bar[c.stage], (it / kStages) & 1u);
        if (kStages == 1 && tid == 0 && t + 1u < tb) {   // single buffer: the next tile waits in L2
            prefetch_l2(W.base + (size_t)(t + 1u) * kTile, tile_bytes_rounded(W, t + 1u));
            if (use_list && listed_count(n_ahead) != 0u)
                prefetch_l2(W.nl_list + (size_t)(t + 1u - W.first_tile) * kNlCap, (listed_count(n_ahead) * 2u + 15u) & ~15u);
        }
        uint32_t n_ahead2 = 0;
        if (use_list && tid == 0) {
            S.tile_total[(it + (uint32_t)kStages) % kRing] = n_ahead;   // read kStages tiles from now
            if (t + (uint32_t)kStages + 1u < tb) n_ahead2 = list_count_of(t + (uint32_t)kStages + 1u);
        }
        NlWords words;
#pragma unroll
        for (int i = 0; i < kWordsPerThread; ++i) words.w[i] = 0u;
        uint32_t total = 0, excl = 0;
        bool listed = false;                           // nlx already holds this tile's newline list
        if (use_list) {
            const uint32_t raw = S.tile_total[it % kRing];   // written kStages tiles ago (or in the prologue)
            total = raw & kNlCountMask;
            listed = list_ok(raw);
        }
        if (listed) {
            const uint16_t* l16 = S.nl16[c.stage];
            for (uint32_t j = tid; j < total; j += kThreads) S.nlx[kHead + j] = c.origin + (uint32_t)l16[j];
            __syncthreads();
        } else {
            if (kCopyStaged && kPack && (kAscii || kQual)) {   // the validation bitmaps reuse `stage`: the previous
                if (tid == 0) tma_store_wait_read();           // tile's bulk stores and edge stores must be through with it
                __syncthreads();
            }
            build_bitmaps<kAscii, kQual>(S, c, addlo, addup);
            __syncthreads();
            words = load_nl_words(S, tid);
            const uint32_t cnt = popc_words(words);
            excl = block_exclusive_scan(S, cnt, total, par);
        }

        // ---- validation from the bitmaps: this thread's 128 bytes, line class known from the rank
        // (a listed tile was screened by k_summarize: no byte of it can fail either validator)
        if ((kAscii || kQual) && !listed) {
            uint32_t r = rank + excl;
#pragma unroll
            for (int i = 0; i < kWordsPerThread; ++i) {
                const uint32_t nl = words.w[i];
                const uint32_t hiw = kAscii ? S.bm_hi_p()[tid * kWordsPerThread + i] : 0u;
                const uint32_t badw = kQual ? (S.bm_bad_p()[tid * kWordsPerThread + i] & ~nl) : 0u;
                if ((hiw | badw) != 0u) {
                    uint32_t rest = 0xFFFFFFFFu, m = nl, rr = r;
                    while (rest) {
                        // segment = bytes up to and including the next newline (or the word's end)
                        uint32_t seg = rest;
                        if (m) { const uint32_t b = __ffs(m) - 1u; seg = rest & (0xFFFFFFFFu >> (31u - b)); m &= m - 1u; }
                        const uint32_t cls = rr & 3u, k = rr >> 2;
                        if (k < n_complete) {
                            if (kAscii && cls != 2u && (hiw & seg)) report(P, k, 4u);
                            if (kQual && cls == 3u && (badw & seg) && P.q5_width == 0u) report(P, k, 5u);
                        }
                        rest &= ~seg;
                        ++rr;
                    }
                }
                r += __popc(nl);
            }
        }

        for (uint32_t pass = 0; pass < total; pass += kNlCap) {
            const uint32_t n = total - pass < (uint32_t)kNlCap ? total - pass : (uint32_t)kNlCap;
            if (!listed) {
                if (total <= (uint32_t)kNlCap) fill_newline_list<false>(S, c, words, excl, 0u);
                else fill_newline_list<true>(S, c, words, excl, pass);
                __syncthreads();
            }
            const uint32_t r0 = rank + pass;  // rank of list entry 0
            const uint32_t d0_id = cum_id, d0_seq = cum_seq, d0_qual = cum_qual;
            // first list entry of each class and the number of lines per class in this pass
            const uint32_t j_id = (0u - r0) & 3u, j_seq = (1u - r0) & 3u, j_qual = (3u - r0) & 3u;
            const uint32_t n_id = j_id < n ? ((n - 1u - j_id) >> 2) + 1u : 0u;
            const uint32_t n_seq = j_seq < n ? ((n - 1u - j_seq) >> 2) + 1u : 0u;
            const uint32_t n_qual = j_qual < n ? ((n - 1u - j_qual) >> 2) + 1u : 0u;
            // does a batch boundary fall among the records that end in this pass?  (uniform)
            const uint32_t k_first = r0 >> 2, k_last = (r0 + n - 1u) >> 2;
            while (edge_k < k_first) edge_k += bsz;
            const bool batch_edge = kPack && edge_k <= k_last;

            // views(): the line-end table, one coalesced store per newline
            if (kOffsets)
                for (uint32_t j = tid; j < n; j += kThreads) P.line_ends[1u + r0 + j] = S.nlx[kHead + j];

            // one thread per RECORD: it handles the (up to four) lines of its record that end in this
            // pass, so every class-specific step runs without divergence and a record's lengths and
            // destinations stay in registers.  Record k owns list entries 4k - r0 + {0,1,2,3}.
            const uint32_t n_rec = k_last - k_first + 1u;
            uint32_t max_len = 0;
            for (uint32_t tb0 = 0; tb0 < n_rec; tb0 += kThreads) {
                const uint32_t tr = tb0 + tid;
                const uint32_t k = k_first + tr;
                const bool mine = tr < n_rec;
                const bool live = mine && k < n_complete;
                const int32_t j0 = (int32_t)(4u * k - r0);                 // list index of the header's newline
                const uint32_t* e = &S.nlx[kHead] + j0;                    // e[c] = newline of class c, e[-1] = the one before
                const bool has0 = mine && j0 >= 0 && j0 < (int32_t)n, has1 = mine && j0 + 1 >= 0 && j0 + 1 < (int32_t)n,
                           has2 = mine && j0 + 2 >= 0 && j0 + 2 < (int32_t)n, has3 = mine && j0 + 3 >= 0 && j0 + 3 < (int32_t)n;
                uint32_t l_id = 0, l_seq = 0, l_qual = 0, s_id = 0, s_seq = 0, s_qual = 0;
                if (has0) {
                    // header line: '@' check (utils.mojo:454), id = line minus '@', stripped (utils.mojo:221-242)
                    const uint32_t q1 = e[-1], p = e[0];
                    const uint32_t len = p - q1 - 1u;
                    uint32_t a = q1 + 2u, nid = 0;
                    if (live) {
                        if (byte_at(S, c, W, q1 + 1u) != '@') report(P, k, 1u);
                        nid = len > 0u ? len - 1u : 0u;
                        if (nid > 0u && (bsq_is_space(byte_at(S, c, W, a)) || bsq_is_space(byte_at(S, c, W, p - 1u)))) {
                            uint32_t z = p;
                            while (a < z && bsq_is_space(byte_at(S, c, W, a))) ++a;
                            while (z > a && bsq_is_space(byte_at(S, c, W, z - 1u))) --z;
                            nid = z - a;
                            if (kPack && P.id_fast) *P.strip_flag = 1u;   // the optimistic id packing is void
                        }
                        if (kOffsets || kPack) { P.id_spans[2u * k] = a; P.id_spans[2u * k + 1u] = nid; }   // (pack: for the strip pipeline)
                        if (len > max_len) max_len = len;
                    }
                    l_id = nid; s_id = a;
                }
                if (has1) {
                    const uint32_t q1 = e[0], len = e[1] - q1 - 1u;
                    if (live) { l_seq = len; bases_acc += len; if (len > max_len) max_len = len; }
                    s_seq = q1 + 1u;
                }
                if (has2 && live && byte_at(S, c, W, e[1] + 1u) != '+') report(P, k, 2u);      // utils.mojo:456
                if (has3) {
                    const uint32_t q1 = e[2], len = e[3] - q1 - 1u;
                    if (live) {
                        if (e[3] - e[-1] > P.rec_limit) report(P, k, 0u);                    // parser.mojo:484-503 (before any other check)
                        if (e[1] - e[0] - 1u != len) report(P, k, 3u);                       // utils.mojo:458-461
                        if (kQual && P.q5_width != 0u && (!listed || q1 + 1u < c.origin)) {
                            // record.mojo:90-102 as written: the first floor(len / W) * W quality bytes fail on
                            // (b - LOWER) >= span, i.e. UPPER itself is rejected there; the rest on > span.
                            // Whether a byte == UPPER counts depends on its offset in the line, so in this mode the
                            // lines of a tile that k_summarize flagged (UPPER counts as suspicious there), and the
                            // lines that began in an earlier tile, are re-examined byte by byte (cold path)
                            const uint32_t body = len - len % P.q5_width;
                            bool bad = false;
                            for (uint32_t x = 0; x < len; ++x) {
                                const uint32_t b = byte_at(S, c, W, q1 + 1u + x);
                                bad = bad || b < P.lower || b > P.upper || (x < body && b == P.upper);
                            }
                            if (bad) report(P, k, 5u);
                        }
                        l_qual = len;
                        if (len > max_len) max_len = len;
                    }
                    s_qual = q1 + 1u;
                }
                if (kPack) {
                    uint32_t e_id, e_seq, e_qual, t_id, t_seq, t_qual;
                    block_exclusive_scan3(S, l_id, l_seq, l_qual, e_id, e_seq, e_qual, t_id, t_seq, t_qual, par);
                    // stream-local line index of this record's lines (class c lines of the pass are consecutive records)
                    if (has0) {
                        const uint32_t i = (uint32_t)(j0 - (int32_t)j_id) >> 2;
                        S.sdst[0][i] = cum_id + e_id + sh_id; S.ssrc[0][i] = s_id;
                        if (live && P.id_fast) {
                            const int64_t endv = P.id_base64 + (int64_t)(cum_id + e_id + l_id);
                            P.id_ends_abs[P.rec_base + (int64_t)k] = endv;
                            if (batch_edge) { const uint32_t tb1 = P.rec_mod + k + 1u;
                                if (tb1 % bsz == 0u) P.id_ends_base[P.rec_div + (int64_t)(tb1 / bsz)] = endv; }
                        }
                    }
                    if (has1) {
                        const uint32_t i = (uint32_t)(j0 + 1 - (int32_t)j_seq) >> 2;
                        S.sdst[1][i] = cum_seq + e_seq + sh_seq; S.ssrc[1][i] = s_seq;
                    }
                    if (has3) {
                        const uint32_t i = (uint32_t)(j0 + 3 - (int32_t)j_qual) >> 2;
                        S.sdst[2][i] = cum_qual + e_qual + sh_qual; S.ssrc[2][i] = s_qual;
                        if (live) {
                            const int64_t endv = P.qual_base64 + (int64_t)(cum_qual + e_qual + l_qual);
                            P.ends_abs[P.rec_base + (int64_t)k] = endv;
                            if (batch_edge) { const uint32_t tb1 = P.rec_mod + k + 1u;
                                if (tb1 % bsz == 0u) P.ends_base[P.rec_div + (int64_t)(tb1 / bsz)] = endv; }
                        }
                    }
                    cum_id += t_id; cum_seq += t_seq; cum_qual += t_qual;
                }
            }
            if (kPack) {
                if (tid < 3) {
                    // two sentinels per stream: the end of the range, and a source for the "next line" read
                    const uint32_t nn = tid == 0 ? n_id : (tid == 1 ? n_seq : n_qual);
                    const uint32_t dd = tid == 0 ? cum_id + sh_id : (tid == 1 ? cum_seq + sh_seq : cum_qual + sh_qual);
                    S.sdst[tid][nn] = dd; S.sdst[tid][nn + 1] = dd; S.sdst[tid][nn + 2] = dd;
                    S.ssrc[tid][nn] = c.origin + 16u; S.ssrc[tid][nn + 1] = c.origin + 16u;
                }
                if (kCopyStaged && tid == 0) tma_store_wait_read();   // the previous pass's bulk stores have read `stage`
                // one barrier: the line tables are complete, every reader of the newline list is done;
                // long lines (long reads, or a line that began far before the tile) are copied
                // vector-parallel from wherever they lie, otherwise straight from [halo | tile]
                const bool long_lines = __syncthreads_or(max_len > (uint32_t)kHalo - 64u) != 0;
                rotate_head(S, n);                     // (thread 0; next read after the barrier that ends the copy)
                // room: bytes the id arena can still take.  Destinations past it only arise after a
                // structure error (an empty header line makes the id prefix diverge); nothing there counts.
                const uint32_t ra_id = d0_id + sh_id, rb_id = cum_id + sh_id;
                const bool do_id = P.id_fast && rb_id > ra_id && (int64_t)rb_id <= P.id_cap - (P.id_base64 - sh_id) && n_id != 0u;
                const uint32_t ra_seq = d0_seq + sh_seq, rb_seq = cum_seq + sh_seq;
                const uint32_t ra_qual = d0_qual + sh_qual, rb_qual = cum_qual + sh_qual;
                const bool ok_seq = rb_seq > ra_seq, ok_qual = rb_qual > ra_qual;
                const uint32_t n1 = ok_seq ? n_seq : 0u, n2 = ok_qual ? n_qual : 0u, n0 = do_id ? n_id : 0u;
                if (long_lines) {
                    StreamJob jobs[3];
                    jobs[0] = StreamJob{S.sdst[0], S.ssrc[0], n_id, ra_id, do_id ? rb_id : ra_id, out_id};
                    jobs[1] = StreamJob{S.sdst[1], S.ssrc[1], n_seq, ra_seq, rb_seq, out_seq};
                    jobs[2] = StreamJob{S.sdst[2], S.ssrc[2], n_qual, ra_qual, rb_qual, out_qual};
#pragma unroll
                    for (int st = 0; st < 3; ++st)
                        if (jobs[st].d1 > jobs[st].d0) copy_stream(S, c, W, jobs[st]);
                    __syncthreads();   // the tile's bytes and the line tables are free again
                } else {
                    const uint8_t* const data = S.data[c.stage];
                    const uint32_t sbias = (uint32_t)kHalo - c.origin;          // window offset -> offset in data[stage]
#if BSQ_COPY_STAGED
                    // stage layout: [id | seq | qual], each region starts at the 16-byte vector of its
                    // first destination byte; lengths are bounded by the bytes of [halo | tile]
                    const uint32_t A_id = ra_id & ~15u, A_seq = ra_seq & ~15u, A_qual = ra_qual & ~15u;
                    const uint32_t so_id = 0u;
                    const uint32_t so_seq = do_id ? ((rb_id - A_id + 15u) & ~15u) : 0u;
                    const uint32_t so_qual = so_seq + (ok_seq ? ((rb_seq - A_seq + 15u) & ~15u) : 0u);
                    // items, one line per thread: sequence and quality lines alternate over the lanes (their
                    // sources start in different banks), then the ids
                    // (only when the distance between consecutive sequence lines is a multiple of 16 bytes, the strides
                    //  whose word loops run in step on few banks: a uniform test on the line table)
                    const bool in_step = BSQ_INTERLEAVE && n1 > 2u && ((S.ssrc[1][1] - S.ssrc[1][0]) & 15u) == 0u &&
                                         ((S.ssrc[1][2] - S.ssrc[1][1]) & 15u) == 0u;
                    const uint32_t npair = in_step ? 2u * (n1 < n2 ? n1 : n2) : 0u, nitem = n1 + n2 + n0;
#if BSQ_ROT == 2
                    // lines a multiple of 16 bytes apart start in few banks: with q = (distance in words) mod 32 (a multiple of
                    // 4 here) a warp's sixteen lines of a stream fall into P = 32 / gcd(q, 32) banks, line i' = lane / 2 of the
                    // warp being the (i' / P)-th on its bank -- that rank is the word its copy starts from
                    const uint32_t qd = in_step ? ((S.ssrc[1][1] - S.ssrc[1][0]) >> 2) & 31u : 4u;
                    const uint32_t lg_p = qd == 0u ? 0u : ((qd & 4u) ? 3u : ((qd & 8u) ? 2u : 1u));
#endif
                    for (uint32_t w0 = 0; w0 < nitem; w0 += kThreads) {
                        const uint32_t w = w0 + tid;
#if BSQ_ROT == 2
                        const uint32_t lanes = (in_step && w < npair) ? ((tid & 31u) >> 1) >> lg_p : 0u;
#else
                        const uint32_t lanes = __ballot_sync(0xFFFFFFFFu, w < nitem);
#endif
                        if (w < nitem) {
                            uint32_t st, i, so, A;
                            if (w < npair) { st = 1u + (w & 1u); i = w >> 1; }
                            else if (!in_step && w < n1 + n2) { st = w < n1 ? 1u : 2u; i = w < n1 ? w : w - n1; }
                            else if (w < n1 + n2) { st = n1 > n2 ? 1u : 2u; i = w - (npair >> 1); }
                            else { st = 0u; i = w - n1 - n2; }
                            if (st == 1u) { so = so_seq; A = A_seq; } else if (st == 2u) { so = so_qual; A = A_qual; } else { so = so_id; A = A_id; }
                            const uint32_t d = S.sdst[st][i];
                            stage_line(data, S.stage, S.ssrc[st][i] + sbias, so + (d - A), S.sdst[st][i + 1] - d, lanes);
                        }
                    }
                    fence_proxy_async_smem();
                    __syncthreads();
                    const uint32_t wsel = tid >> 5;
                    if (do_id) flush_stream(S.stage, so_id, ra_id, rb_id, out_id, tid == 0, wsel == 0u);
                    if (ok_seq) flush_stream(S.stage, so_seq, ra_seq, rb_seq, out_seq, tid == 0, wsel == 1u);
                    if (ok_qual) flush_stream(S.stage, so_qual, ra_qual, rb_qual, out_qual, tid == 0, wsel == 2u);
                    if (tid == 0) tma_store_commit();
#else
                    const DirectJob jq{S.sdst[1], S.ssrc[1], n_seq, ra_seq, rb_seq, out_seq};
                    const DirectJob jr{S.sdst[2], S.ssrc[2], n_qual, ra_qual, rb_qual, out_qual};
                    if (ok_seq) copy_interior(data, sbias, jq);
                    if (ok_qual) copy_interior(data, sbias, jr);
                    // one thread per line: the edge vectors of the sequence and quality lines, then the ids
                    for (uint32_t w = tid; w < n1 + n2 + n0; w += kThreads) {
                        if (w < n1) copy_edges(data, sbias, jq, w);
                        else if (w < n1 + n2) copy_edges(data, sbias, jr, w - n1);
                        else {
                            const uint32_t i = w - n1 - n2, d = S.sdst[0][i];
                            copy_small_line(data, S.ssrc[0][i] + sbias, out_id, d, S.sdst[0][i + 1] - d);
                        }
                    }
                    __syncthreads();   // the tile's bytes and the line tables are free again
#endif
                }
            } else {
                __syncthreads();
                rotate_head(S, n);
                if (pass + (uint32_t)kNlCap < total || use_list) __syncthreads();   // the next pass / tile refills the list
            }
        }
        rank += total;
        if (total == 0u) __syncthreads();   // (the pass loop ends with a barrier otherwise)
        if (tid == 0 && t + (uint32_t)kStages < tb)
            issue_tile_load<true>(S, W, t + (uint32_t)kStages, c.stage, S.nl16[c.stage], use_list ? listed_count(n_ahead) : 0u);
        n_ahead = n_ahead2;
    }
    if (kCopyStaged && kPack && tid == 0) tma_store_wait_all();   // shared memory must outlive the bulk stores
    // one atomic per warp: the run's share of the base count
    unsigned long long b64 = bases_acc;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) b64 += __shfl_xor_sync(0xFFFFFFFFu, b64, d);
    if ((tid & 31u) == 0u && b64 != 0ull) atomicAdd(P.bases, b64);
}

}  // namespace bsq
