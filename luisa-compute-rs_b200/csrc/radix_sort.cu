// radix_sort.cu — onesweep LSD radix sort of (uint64 key, uint32 value) pairs for the LBVH builder.
//
// One histogram pre-pass counts all digits of all passes; every pass is then a single kernel
// ("one sweep" over the data): tiles take dynamic tickets, rank their keys with warp-level
// __match_any_sync multi-split, and obtain their global digit offsets through a chained scan
// with decoupled look-back (Adinets & Merrill 2022; Merrill & Garland 2016).  Stable.
//
// Each tile is reordered by digit in shared memory before it is written, so every digit's keys leave as one
// contiguous run (full sectors) instead of 8-byte scattered stores.
// HBM traffic per pass: read 12 B + write 12 B per pair.  Histogram pre-pass: read 8 B per key.
#include "build.cuh"

namespace lcb {

namespace {

constexpr int kRadix = 256;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
#ifndef LCB_SORT_ITEMS_SMALL
#define LCB_SORT_ITEMS_SMALL 16
#endif
constexpr int kItemsLarge = 16, kItemsSmall = LCB_SORT_ITEMS_SMALL;   // keys per thread: tiles of 4096 pairs (8 and 12 per thread measured slower at 1 M pairs: profiles/r02n_variants.txt)
constexpr uint32_t kSmallTileMax = 4u << 20;      // sorts up to this size take the small tile
constexpr uint32_t kFlagAgg = 1u << 30, kFlagPrefix = 2u << 30, kValueMask = (1u << 30) - 1;
constexpr int kSmallSortMax = 2048;
constexpr int kLookBatch = 8;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

// Histogram of every 8-bit digit of every pass in one read of the keys.
template <typename K>
__global__ void __launch_bounds__(256) k_sort_histogram(const K *__restrict__ keys, uint32_t n, uint32_t *__restrict__ ghist,
                                                         int begin_bit, int passes) {
    __shared__ uint32_t sh[8 * kRadix];
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        K k = keys[i] >> begin_bit;
        for (int p = 0; p < passes; p++) {
            atomicAdd(&sh[p * kRadix + (uint32_t)(k & 0xff)], 1u);
            k >>= 8;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        uint32_t c = sh[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// Exclusive scan of each pass's 256 bins (one block per pass).
__global__ void __launch_bounds__(256) k_sort_scan(uint32_t *ghist) {
    __shared__ uint32_t sh[kRadix];
    uint32_t *h = ghist + blockIdx.x * kRadix;
    uint32_t v = h[threadIdx.x];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kRadix; off <<= 1) {
        uint32_t t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    h[threadIdx.x] = sh[threadIdx.x] - v;
}

// K: key type (uint32_t when the Morton code fits — up to 2^25 primitives — else uint64_t); kItems: keys per thread.  Small sorts take
// smaller tiles: with 16 keys per thread a 1 M-pair pass is 245 tiles on 148 SMs, one serial chain of load / rank / look-back / scatter
// per tile with little to overlap it; the small tile is sized so that a 1 M-pair pass still fits one wave of resident CTAs.
template <typename K, int kItems>
__global__ void __launch_bounds__(kSortThreads, kItems * sizeof(K) <= 64 ? 3 : 2) k_sort_onesweep(const K *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                                K *__restrict__ kout, uint32_t *__restrict__ vout, uint32_t n, int shift,
                                                                const uint32_t *__restrict__ gbase, uint32_t *status, uint32_t *ticket) {
    constexpr int kTile = kSortThreads * kItems;
    constexpr int kEntry = sizeof(K) > 5 ? sizeof(K) : 5;   // the exchange buffer holds the keys, then values (4 B) + their digits (1 B)
    __shared__ uint32_t warp_hist[kSortWarps][kRadix];
    __shared__ uint32_t tile_start[kRadix];   // first position of digit d inside the (digit-sorted) tile
    __shared__ int digit_delta[kRadix];       // global position of digit d's run minus its position inside the tile
    __shared__ __align__(8) uint8_t exch_raw[kTile * kEntry];  // tile in digit order: keys, then (reused) values -> coalesced runs per digit
    K *exch = reinterpret_cast<K *>(exch_raw);
    __shared__ uint32_t tile_s;
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = tile_s;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1;
    const uint32_t base = tile * kTile + warp * (32 * kItems);
    const uint32_t tile_count = min((uint32_t)kTile, n - tile * kTile);

    K key[kItems];
    uint16_t rank[kItems];
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        uint32_t idx = base + j * 32 + lane;
        key[j] = idx < n ? kin[idx] : (K)~(K)0;
    }
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        uint32_t idx = base + j * 32 + lane;
        bool valid = idx < n;
        uint32_t d = (uint32_t)(key[j] >> shift) & 0xff;
        uint32_t m = __match_any_sync(0xffffffffu, valid ? d : (0x100u + lane));
        int leader = __ffs(m) - 1;
        uint32_t pre = 0;
        if (valid && (int)lane == leader) {
            pre = warp_hist[warp][d];
            warp_hist[warp][d] = pre + __popc(m);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[j] = (uint16_t)(pre + __popc(m & lt_mask));
        __syncwarp();
    }
    __syncthreads();

    uint32_t total;
    {   // thread d owns digit d: warp offsets inside the tile, exclusive scan of the tile's digit counts, decoupled look-back
        const uint32_t d = threadIdx.x;
        total = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) { uint32_t c = warp_hist[w][d]; warp_hist[w][d] = total; total += c; }
        // block-wide exclusive scan of `total` over the 256 digits (warp shuffles + one shared round)
        uint32_t incl = total;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= (uint32_t)off) incl += t; }
        __shared__ uint32_t warp_sum[kSortWarps];
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        uint32_t warp_base = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) if ((uint32_t)w < warp) warp_base += warp_sum[w];
        const uint32_t start = warp_base + incl - total;
        tile_start[d] = start;
        uint32_t *mine = status + (size_t)tile * kRadix + d;
        uint32_t excl = 0;
        if (tile == 0) {
            st_volatile_u32(mine, kFlagPrefix | total);
        } else {
            st_volatile_u32(mine, kFlagAgg | total);
            // look-back, kLookBatch predecessors per L2 round trip: when every tile of a pass is resident at once (1 M pairs = 245 tiles
            // on 148 SMs) nobody has an inclusive prefix to offer at first, and a one-at-a-time walk is a chain of ~sqrt(2 * tiles)
            // dependent loads per pass; the batch is consumed in order and stops at the first prefix or the first unpublished entry
            int look = (int)tile - 1;
            bool found = false;
            while (!found) {
                uint32_t v[kLookBatch];
#pragma unroll
                for (int b = 0; b < kLookBatch; b++) v[b] = look - b >= 0 ? ld_volatile_u32(status + (size_t)(look - b) * kRadix + d) : 0u;
#pragma unroll
                for (int b = 0; b < kLookBatch; b++) {
                    if (found) break;
                    const uint32_t f = v[b] >> 30;
                    if (f == 0) break;            // not published yet: poll again from here
                    excl += v[b] & kValueMask;
                    look--;
                    if (f == 2) found = true;
                }
            }
            st_volatile_u32(mine, kFlagPrefix | (excl + total));
        }
        digit_delta[d] = (int)(gbase[d] + excl) - (int)start;
    }
    __syncthreads();
    // keys: registers -> shared in digit order -> global in runs
    uint16_t pos[kItems];
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        uint32_t idx = base + j * 32 + lane;
        uint32_t d = (uint32_t)(key[j] >> shift) & 0xff;
        pos[j] = (uint16_t)(tile_start[d] + warp_hist[warp][d] + rank[j]);
        if (idx < n) exch[pos[j]] = key[j];
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < tile_count; p += kSortThreads) {
        const K k = exch[p];
        const uint32_t d = (uint32_t)(k >> shift) & 0xff;
        kout[(int)p + digit_delta[d]] = k;
    }
    __syncthreads();
    // values: same route through the (reused) exchange buffer; the digit of position p is recovered from the key pass
    uint32_t *exv = reinterpret_cast<uint32_t *>(exch_raw);
    uint8_t *exd = exch_raw + kTile * 4;
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        uint32_t idx = base + j * 32 + lane;
        if (idx < n) { exv[pos[j]] = vin[idx]; exd[pos[j]] = (uint8_t)((key[j] >> shift) & 0xff); }
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < tile_count; p += kSortThreads) vout[(int)p + digit_delta[exd[p]]] = exv[p];
}

// n <= 2048: one block, bitonic sort of (key, value) in shared memory; (key, value) compared
// lexicographically so the result equals the stable sort when values are the input positions.
template <typename K>
__global__ void __launch_bounds__(1024) k_sort_small(K *keys, uint32_t *vals, uint32_t n) {
    __shared__ K sk[kSmallSortMax];
    __shared__ uint32_t sv[kSmallSortMax];
    for (uint32_t i = threadIdx.x; i < kSmallSortMax; i += blockDim.x) {
        sk[i] = i < n ? keys[i] : (K)~(K)0;
        sv[i] = i < n ? vals[i] : 0xffffffffu;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= kSmallSortMax; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < kSmallSortMax; i += blockDim.x) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    bool up = (i & k) == 0;
                    K a = sk[i], b = sk[ixj];
                    uint32_t va = sv[i], vb = sv[ixj];
                    bool gt = a > b || (a == b && va > vb);
                    if (gt == up) { sk[i] = b; sk[ixj] = a; sv[i] = vb; sv[ixj] = va; }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { keys[i] = sk[i]; vals[i] = sv[i]; }
}

}  // namespace

static int tile_pairs(uint32_t n) { return kSortThreads * (n <= kSmallTileMax ? kItemsSmall : kItemsLarge); }

size_t sort_scratch_bytes(uint32_t n, int passes) {
    if (n <= (uint32_t)kSmallSortMax) return 256;
    size_t tiles = (n + tile_pairs(n) - 1) / tile_pairs(n);
    return (size_t)8 * kRadix * 4 + 256 + (size_t)passes * tiles * kRadix * 4;
}

template <typename K>
static bool sort_pairs_t(cudaStream_t s, uint32_t n, K *keys, uint32_t *vals, K *keys_alt, uint32_t *vals_alt, void *scratch,
                         int begin_bit, int passes, LaunchCounter &lc) {
    if (n <= (uint32_t)kSmallSortMax) {
        k_sort_small<K><<<1, 1024, 0, s>>>(keys, vals, n); lc.count++;
        return false;
    }
    const bool small = n <= kSmallTileMax;
    const size_t tiles = (n + tile_pairs(n) - 1) / tile_pairs(n);
    uint32_t *ghist = (uint32_t *)scratch;
    uint32_t *tickets = ghist + 8 * kRadix;
    uint32_t *status = tickets + 64;
    cudaMemsetAsync(scratch, 0, sort_scratch_bytes(n, passes), s);
    int hist_blocks = (int)((n + 256 * 16 - 1) / (256 * 16));
    if (hist_blocks > 148 * 8) hist_blocks = 148 * 8;
    k_sort_histogram<K><<<hist_blocks, 256, 0, s>>>(keys, n, ghist, begin_bit, passes); lc.count++;
    k_sort_scan<<<passes, kRadix, 0, s>>>(ghist); lc.count++;
    K *ki = keys, *ko = keys_alt;
    uint32_t *vi = vals, *vo = vals_alt;
    for (int p = 0; p < passes; p++) {
        if (small) k_sort_onesweep<K, kItemsSmall><<<(unsigned)tiles, kSortThreads, 0, s>>>(ki, vi, ko, vo, n, begin_bit + 8 * p, ghist + p * kRadix, status + (size_t)p * tiles * kRadix, tickets + p);
        else k_sort_onesweep<K, kItemsLarge><<<(unsigned)tiles, kSortThreads, 0, s>>>(ki, vi, ko, vo, n, begin_bit + 8 * p, ghist + p * kRadix, status + (size_t)p * tiles * kRadix, tickets + p);
        lc.count++;
        K *tk = ki; ki = ko; ko = tk;
        uint32_t *tv = vi; vi = vo; vo = tv;
    }
    return (passes & 1) != 0;
}

// Sorts by bits [begin_bit, begin_bit + 8*passes).  Returns true if the result is in (keys_alt, vals_alt).  key_bytes = 4: the key
// arrays hold uint32_t (the caller's Morton codes fit 32 bits), 8: uint64_t.
bool sort_pairs(cudaStream_t s, uint32_t n, void *keys, uint32_t *vals, void *keys_alt, uint32_t *vals_alt, void *scratch,
                int begin_bit, int passes, int key_bytes, LaunchCounter &lc) {
    if (n <= 1) return false;
    if (key_bytes == 4) return sort_pairs_t<uint32_t>(s, n, (uint32_t *)keys, vals, (uint32_t *)keys_alt, vals_alt, scratch, begin_bit, passes, lc);
    return sort_pairs_t<uint64_t>(s, n, (uint64_t *)keys, vals, (uint64_t *)keys_alt, vals_alt, scratch, begin_bit, passes, lc);
}

}  // namespace lcb
