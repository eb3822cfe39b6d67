// Declarations shared by the two SPIHT encoder kernels: spiht_enc.cu (one CTA per image) and spiht_enc_cl.cu
// (one thread-block cluster per image).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace spihtb {

constexpr int ENC_NT = 512;
constexpr int ENC_ITEMS = 4;                          // consecutive list entries per thread and scan
constexpr int ENC_CHUNK = ENC_NT * ENC_ITEMS;
constexpr int ENC_REF_ITEMS = 8;                      // refinement: entries per thread per flush
constexpr int ENC_RING = 1024;                        // staging ring in words; a chunk emits at most 9 bits per entry
constexpr int ENC_SLACK = 4 * ENC_CHUNK + 64;         // list slack for the chunk that crosses the budget
static_assert(ENC_CHUNK * 9 / 32 + 8 < ENC_RING, "staging ring too small");
static_assert((ENC_RING & (ENC_RING - 1)) == 0, "ring size must be a power of two");
static_assert(ENC_CHUNK <= 2048, "work-list words pack a chunk index into 11 bits and a rank into 12");

struct EncK {
    const int32_t *coeffs;
    int B, C, H, W, NH, NW, ll_h, ll_w;
    KeyFmt kf;
    const uint8_t *dp, *lp, *dpll, *lpll;
    const uint32_t *maxabs;
    uint64_t max_bits;
    const uint64_t *dev_max_bits;
    uint32_t *out;
    uint64_t out_stride_words;
    uint64_t *nbits;
    int32_t *max_n;
    int32_t *status;
    // per-slot list storage
    int32_t *lip;
    uint32_t *lsp;
    uint2 *lis;  // 3 buffers per slot: R, G0, G1
    size_t pix_cap, lis_cap;
    unsigned int *counter;
};

// ---- gathers through cp.async (LDGSTS): 4-byte copies global -> shared
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void *gptr)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// the aligned 32-bit word holding byte *p, and that byte's value within the word
__device__ __forceinline__ const void *word_of(const uint8_t *p)
{
    return reinterpret_cast<const void *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
}
__device__ __forceinline__ uint32_t byte_of(uint32_t word, const uint8_t *p)
{
    return (word >> (8u * (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u))) & 0xffu;
}
__device__ __forceinline__ int4 lds_int4(uint32_t saddr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// OR `nb` (<= 40) bits of `val` into the staging ring at stream position `off`; bits at or past `limit` are dropped
__device__ __forceinline__ void bw_emit(uint32_t *ring, uint64_t limit, uint64_t off, uint64_t val, int nb)
{
    if (nb == 0 || off >= limit) return;
    if (off + (uint64_t)nb > limit) {
        nb = (int)(limit - off);
        val &= (1ull << nb) - 1ull;  // here 1 <= nb < 40
    }
    const uint32_t w = (uint32_t)(off >> 5);
    const int sh = (int)(off & 31);
    const uint32_t first = (uint32_t)(val << sh);
    if (first) atomicOr(&ring[w & (ENC_RING - 1)], first);
    const uint64_t rest = sh ? (val >> (32 - sh)) : (val >> 32);
    if (rest) {
        atomicOr(&ring[(w + 1) & (ENC_RING - 1)], (uint32_t)rest);
        if (rest >> 32) atomicOr(&ring[(w + 2) & (ENC_RING - 1)], (uint32_t)(rest >> 32));
    }
}

// Write out (and clear) the complete words below stream position `end`.  The partial word stays.  The
// next emits only touch words at or above it, and a cleared slot is reused no earlier than a whole ring
// later, which is at least one barrier away.
__device__ __forceinline__ void bw_flush(uint32_t *ring, uint32_t *outrow, uint64_t &wflushed, uint64_t end)
{
    __syncthreads();
    const uint64_t wend = end >> 5;
    for (uint64_t w = wflushed + threadIdx.x; w < wend; w += ENC_NT) {
        outrow[w] = ring[w & (ENC_RING - 1)];
        ring[w & (ENC_RING - 1)] = 0u;
    }
    wflushed = wend;
}


}  // namespace spihtb
