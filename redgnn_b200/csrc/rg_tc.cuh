// tcgen05 / TMEM / mbarrier PTX wrappers shared by the tensor-core node kernels (sm_100a only).
#pragma once
#include "rg_common.cuh"

namespace rgtc {

constexpr int kTcRows = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T ; one K=8 TF32 slice
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// store 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :
        : "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
          "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
          "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor: K-major, no swizzle (INTERLEAVE), sm_100 version bits
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// TF32 split x = hi + lo with hi ROUNDED TO NEAREST on 11 significant bits (not truncated): |lo| <= 2^-12 |x|,
// so lo has at most 12 significant bits and the tensor core's own truncation of lo to 11 drops one bit
// (2^-23 |x|) instead of two to three (2^-21.4 |x|) -- measured on the power-law shape (hub rows with
// |agg| ~ 10^3): score error vs fp64 2.1e-4 -> see DESIGN.md 4.4.
__device__ __forceinline__ float tf32_hi(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x00001000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float4 x, float4 &hi, float4 &lo) {
    hi.x = tf32_hi(x.x);
    hi.y = tf32_hi(x.y);
    hi.z = tf32_hi(x.z);
    hi.w = tf32_hi(x.w);
    lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}

// "Lane-interleaved" plane layout of the buffers that only the tensor-core node kernels exchange (the
// saved gate planes, G4, g_pre): rows in tiles of 32, inside a tile chunk-major,
//     [tile = row / 32][chunk = col / 4][row % 32][4 floats]  (+ 4 pad floats per chunk).
// In those kernels lane == node row (TMEM lane), so with row-major rows of 4*D bytes every 128-bit load /
// store instruction of a warp touched 32 different cache lines (ncu: the LSU wavefront pipe was the top
// limiter of k_node_bwd_tc at 56 %); here the 32 lanes of one instruction cover 512 contiguous bytes.
// A 32-row tile is still ONE contiguous run of kIlTile(D) floats (what the weight-gradient kernel
// bulk-copies into shared memory); the 16-byte pad per chunk staggers the chunks over the shared-memory
// banks there (unpadded, all chunks of a row sit 512 B apart = on the same banks).
constexpr int kIlChunk = 132;                                                     // floats per chunk of a tile
__host__ __device__ constexpr int il_tile_floats(int D) { return (D / 4) * kIlChunk; }
__host__ __device__ __forceinline__ size_t il_plane_floats(int64_t rows, int D) {
    return (size_t)((rows + 31) >> 5) * il_tile_floats(D);
}
__device__ __forceinline__ size_t il_off(int64_t row, int chunk, int chunks_per_row) {
    return ((size_t)(row >> 5) * chunks_per_row + chunk) * kIlChunk + (size_t)(row & 31) * 4;   // in floats
}

// D[tmem] (+)= A[tmem: lane = row, 8 consecutive 32-bit columns = K] . B[smem]^T ; one K=8 TF32 slice
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

}  // namespace rgtc
