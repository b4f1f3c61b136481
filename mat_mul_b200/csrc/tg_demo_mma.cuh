// tg_demo_mma.cuh -- the per-warp routines of K3m (see tg_demo_mma.cu): action records -> halves in shared memory,
// then the 16 GEMMs of one 16x16x16 demo on mma.sync f16.  Shared by the stand-alone accumulation kernel
// (tg_demo_mma.cu) and the fused generation kernel (tg_demo.cu).
#pragma once
#include <cuda_fp16.h>

#include "tg_common.cuh"

namespace tg {
namespace acc16 {

constexpr int HP = 28;                    // uint32 per H row: 24 + 4 (ldmatrix rows 112 bytes apart: conflict-free)
constexpr int H_WORDS = 64 * HP;          // 1792 words = 7 KB (the 4 KB output tile reuses it)
constexpr int UP = 36;                    // uint32 per U2 row: 32 + 4
constexpr int U_WORDS = 16 * UP;          // 576 words
constexpr int WARP_WORDS = H_WORDS + U_WORDS;
constexpr float BIAS = 12582912.0f + 64.0f; // 1.5 * 2^23 + 64

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hsub2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// four 8x8 b16 matrices, transposed: lane (g,t) gets (row 2t, col g | row 2t+1, col g) of each
__device__ __forceinline__ void ldsm4t(uint32_t (&d)[4], const void *row_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(smem_u32(row_ptr)));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}


// Step 1 for one record: 48 token bytes (three uint4) -> row r of H; returns bit 7 set in some byte iff a token is
// outside [0, 2 shift]
__device__ __forceinline__ uint32_t record_to_halves(const uint4 (&raw)[3], int shift, uint32_t *sH_row) {
    const uint32_t c64 = 0x64646464u;
    const __half2 off2 = __floats2half2_rn(1024.0f + (float)shift, 1024.0f + (float)shift);
    const uint32_t off = *reinterpret_cast<const uint32_t *>(&off2);
    const uint32_t vadd = (uint32_t)(0x7F - 2 * shift) * ONES4; // byte + vadd sets bit 7 iff byte > 2 shift (bytes < 128)
    uint32_t w[12], h[24], invalid = 0;
#pragma unroll
    for (int q = 0; q < 3; q++) w[4 * q] = raw[q].x, w[4 * q + 1] = raw[q].y, w[4 * q + 2] = raw[q].z, w[4 * q + 3] = raw[q].w;
#pragma unroll
    for (int q = 0; q < 12; q++) {
        invalid |= (((w[q] & 0x7F7F7F7Fu) + vadd) | w[q]) & H4;
        h[2 * q] = hsub2(prmt(w[q], c64, 0x4140u), off);
        h[2 * q + 1] = hsub2(prmt(w[q], c64, 0x4342u), off);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(sH_row);
#pragma unroll
    for (int q = 0; q < 6; q++) dst[q] = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
    return invalid;
}

// Steps 2 and 3: H (rows 0 .. 16 KS - 1 written, visible to the warp) -> the demo's 4096 bytes at `out` (one TMA bulk
// store, waited for); returns true iff some entry is outside [-64, 63].  sH is reused for the output tile.
template <int KS, int UNR = 4>
__device__ __forceinline__ bool gemms_from_halves(uint32_t *sH, uint32_t *sU, int8_t *__restrict__ out, int lane) {
    const int g = lane >> 2, t = lane & 3;
    // ---------------- 2. fragments: lane l addresses row (l & 7) of matrix l >> 3
    const int mi = lane >> 3, rr = lane & 7;
    uint32_t vf[KS][4], wf[KS][4];
#pragma unroll
    for (int s = 0; s < KS; s++) {
        // A = V^T: matrices (r lo, j lo), (r lo, j hi), (r hi, j lo), (r hi, j hi) -> a0, a1, a2, a3
        ldsm4t(vf[s], sH + (16 * s + 8 * (mi >> 1) + rr) * HP + 8 + 4 * (mi & 1));
        // B = W: matrices (r lo, k lo), (r hi, k lo), (r lo, k hi), (r hi, k hi) -> b0, b1 of tile 0, b0, b1 of tile 1
        ldsm4t(wf[s], sH + (16 * s + 8 * (mi & 1) + rr) * HP + 16 + 4 * (mi >> 1));
    }
    {
        // U pairs: matrices (r of slot q, i lo), (same, i hi) for two slots per ldmatrix; slot q = 2s + hi  <->  r = 16s + 8hi + 2t
        uint32_t ulo[2 * KS], uhi[2 * KS]; // i = g | g + 8
#pragma unroll
        for (int q2 = 0; q2 < KS; q2++) {
            uint32_t d[4];
            ldsm4t(d, sH + (8 * (2 * q2 + (mi >> 1)) + rr) * HP + 4 * (mi & 1));
            ulo[2 * q2] = d[0], uhi[2 * q2] = d[1], ulo[2 * q2 + 1] = d[2], uhi[2 * q2 + 1] = d[3];
        }
        uint32_t *d0 = sU + g * UP + 8 * t, *d1 = sU + (g + 8) * UP + 8 * t;
#pragma unroll
        for (int q = 0; q < 2 * KS; q += 2) {
            *reinterpret_cast<uint2 *>(d0 + q) = make_uint2(ulo[q], ulo[q + 1]);
            *reinterpret_cast<uint2 *>(d1 + q) = make_uint2(uhi[q], uhi[q + 1]);
        }
    }
    __syncwarp(); // H is dead from here on: the output tile takes its place

    // ---------------- 3. one GEMM per i
    uint8_t *tile = reinterpret_cast<uint8_t *>(sH);
    uint32_t chk = 0;
#pragma unroll UNR
    for (int i = 0; i < 16; i++) {
        uint32_t ub[2 * KS];
#pragma unroll
        for (int q = 0; q < 2 * KS; q += 2) {
            const uint2 u2 = *reinterpret_cast<const uint2 *>(sU + i * UP + 8 * t + q);
            ub[q] = u2.x, ub[q + 1] = u2.y;
        }
        float acc0[4] = {BIAS, BIAS, BIAS, BIAS}, acc1[4] = {BIAS, BIAS, BIAS, BIAS}; // k = 2t, 2t+1 | 8+2t, 9+2t; rows j = g | g+8
#pragma unroll
        for (int s = 0; s < KS; s++) {
            uint32_t a[4];
            a[0] = hmul2(vf[s][0], ub[2 * s]), a[1] = hmul2(vf[s][1], ub[2 * s]);
            a[2] = hmul2(vf[s][2], ub[2 * s + 1]), a[3] = hmul2(vf[s][3], ub[2 * s + 1]);
            mma_f16(acc0, a, wf[s][0], wf[s][1]);
            mma_f16(acc1, a, wf[s][2], wf[s][3]);
        }
        // T + 64 in [0, 127]  <=>  mantissa bits 7..22 equal those of 1.5 * 2^23
        chk |= (__float_as_uint(acc0[0]) | __float_as_uint(acc0[1])) | (__float_as_uint(acc0[2]) | __float_as_uint(acc0[3]));
        chk |= (__float_as_uint(acc1[0]) | __float_as_uint(acc1[1])) | (__float_as_uint(acc1[2]) | __float_as_uint(acc1[3]));
#pragma unroll
        for (int x = 0; x < 4; x++) acc0[x] -= 64.0f, acc1[x] -= 64.0f;
        uint8_t *row = tile + (i * 16 + g) * 16 + 2 * t;
        *reinterpret_cast<uint16_t *>(row) = (uint16_t)prmt(__float_as_uint(acc0[0]), __float_as_uint(acc0[1]), 0x0040u);
        *reinterpret_cast<uint16_t *>(row + 8) = (uint16_t)prmt(__float_as_uint(acc1[0]), __float_as_uint(acc1[1]), 0x0040u);
        *reinterpret_cast<uint16_t *>(row + 128) = (uint16_t)prmt(__float_as_uint(acc0[2]), __float_as_uint(acc0[3]), 0x0040u);
        *reinterpret_cast<uint16_t *>(row + 136) = (uint16_t)prmt(__float_as_uint(acc1[2]), __float_as_uint(acc1[3]), 0x0040u);
    }
    // every register was 0x4B400000 + (T + 64): any bit of 7..21 set, or bit 22 cleared (negative), means out of range
    const bool bad = __any_sync(0xFFFFFFFFu, ((chk ^ 0x4B400000u) & 0xFFFFFF80u) != 0);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        bulk_s2g(out, tile, 4096u);
        bulk_commit();
        bulk_wait<0>();
    }
    __syncwarp(); // the tile (= H) may be rewritten once lane 0 is through
    return bad;
}

} // namespace acc16
} // namespace tg
