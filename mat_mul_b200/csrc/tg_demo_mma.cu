// tg_demo_mma.cu -- K3m: the target tensor of a 16x16x16 action list on the tensor cores (mma.sync f16), used by
// tg_demo_accumulate and tg_demo_gen_philox for S = 16, R <= 64.
//
// Reference restated: target = sum_r u_r (x) v_r (x) w_r (utils.py:40-53 uvw_to_demo, utils.py:232,
// datasets.py:141).  Per demo and per first index i this is one GEMM
//     T[i][j][k] = sum_r (u_r[i] v_r[j]) * w_r[k]          M = j (16), N = k (16), K = r (padded to 16 KS)
// in f16 with f32 accumulation -- exact integer arithmetic: coefficients |c| <= 4, |u v| <= 16, |T| <= 64 R.
// One WARP per demo:
//   1. lane l converts action records l and l + 32 (48 token bytes each, straight from the step-major tape) into
//      halves, two tokens per PRMT (byte b -> half 1024 + b) + one HSUB2, rows H[r][48] in shared memory;
//   2. ldmatrix.trans turns H into MMA fragments whose K slots run along r without any explicit transposition:
//      V^T as the A operand (16 registers), W as the B operand (16 registers), and the pairs (u_2p[i], u_2p+1[i]) of
//      U, which go back to shared memory as the broadcast table U2[i][t][8];
//   3. for every i: A = V-fragment (.) broadcast(U2[i]) -- the Khatri-Rao rows u_r[i] v_r[j], ONE HMUL2 per fragment
//      register -- and 2 KS HMMAs.  The accumulators start at 1.5 * 2^23 + 64, so their bit patterns hold T + 64 as
//      integers: the range test [-64, 63] is an OR over the raw registers, and after one FADD the low byte is the
//      int8 entry.  Rows leave through a 4 KB tile (over H) and one TMA bulk store per demo.
// A demo with a token outside [0, 2 shift] (not an action of this alphabet; its coefficient is int8(token - shift) in
// the packed-IMAD kernel) is summed entry by entry in int32 instead -- identical results.
// Measured against the packed-IMAD kernel of tg_demo.cu and the tcgen05 kernel of tg_demo_tc.cu: profiles/README.md.
#include <cuda_fp16.h>

#include <cstdlib>

#include "tg_common.cuh"

namespace tg {

namespace {

constexpr int WARPS = 4;                  // demos per CTA
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

// KS = ceil(R / 16) K-steps.  or_flags: the flags array already holds TG_FLAG_EXHAUSTED bits of the sampler.
template <int KS>
__global__ void __launch_bounds__(32 * WARPS)
    demo_acc16_mma_kernel(const uint8_t *__restrict__ tape, long long tape_step_stride, long long N, int R, int shift,
                          int8_t *__restrict__ slab, uint8_t *__restrict__ flags, int or_flags) {
    extern __shared__ __align__(128) uint32_t s_words[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const long long n = (long long)blockIdx.x * WARPS + warp;
    if (n >= N) return;
    uint32_t *sH = s_words + warp * WARP_WORDS, *sU = sH + H_WORDS;

    // ---------------- 1. records -> halves
    const uint32_t c64 = 0x64646464u;
    const __half2 off2 = __floats2half2_rn(1024.0f + (float)shift, 1024.0f + (float)shift);
    const uint32_t off = *reinterpret_cast<const uint32_t *>(&off2);
    const uint32_t vadd = (uint32_t)(0x7F - 2 * shift) * ONES4; // byte + vadd sets bit 7 iff byte > 2 shift (bytes < 128)
    uint32_t invalid = 0;
#pragma unroll
    for (int pass = 0; pass < (KS + 1) / 2; pass++) {
        const int r = 32 * pass + lane;
        if (r < 16 * KS) {
            uint32_t w[12];
            if (r < R) {
                const uint4 *src = reinterpret_cast<const uint4 *>(tape + (size_t)r * tape_step_stride + n * 48);
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const uint4 v4 = __ldg(src + q);
                    w[4 * q] = v4.x, w[4 * q + 1] = v4.y, w[4 * q + 2] = v4.z, w[4 * q + 3] = v4.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 12; q++) w[q] = (uint32_t)shift * ONES4; // coefficient 0
            }
            uint32_t h[24];
#pragma unroll
            for (int q = 0; q < 12; q++) {
                invalid |= (((w[q] & 0x7F7F7F7Fu) + vadd) | w[q]) & H4;
                h[2 * q] = hsub2(prmt(w[q], c64, 0x4140u), off);
                h[2 * q + 1] = hsub2(prmt(w[q], c64, 0x4342u), off);
            }
            uint4 *dst = reinterpret_cast<uint4 *>(sH + r * HP);
#pragma unroll
            for (int q = 0; q < 6; q++) dst[q] = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        }
    }
    int8_t *out = slab + n * 4096;
    if (__any_sync(0xFFFFFFFFu, invalid != 0)) {
        // not an action list of this alphabet: entry by entry in int32, coefficient = int8(token - shift)
        uint32_t bad = 0;
        for (int e = lane; e < 4096; e += 32) {
            const int i = e >> 8, j = (e >> 4) & 15, k = e & 15;
            int acc = 0;
            for (int r = 0; r < R; r++) {
                const uint8_t *rec = tape + (size_t)r * tape_step_stride + n * 48;
                acc += (int)(int8_t)(rec[i] - shift) * (int)(int8_t)(rec[16 + j] - shift) * (int)(int8_t)(rec[32 + k] - shift);
            }
            if (acc < -64 || acc > 63) bad = 1;
            out[e] = (int8_t)acc;
        }
        bad = __any_sync(0xFFFFFFFFu, bad != 0);
        if (flags && lane == 0) flags[n] = (uint8_t)((or_flags ? flags[n] : 0) | (bad ? TG_FLAG_RANGE : 0));
        return;
    }
    __syncwarp();

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
#pragma unroll 2
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
        if (flags) flags[n] = (uint8_t)((or_flags ? flags[n] : 0) | (bad ? TG_FLAG_RANGE : 0));
        bulk_wait<0>();
    }
}

} // namespace

bool demo_acc16_mma_applies(int R) { return R >= 1 && R <= 64; }

int launch_demo_acc16_mma(const uint8_t *tape, long long tape_step_stride, long long N, int R, int shift, int8_t *slab,
                          uint8_t *flags, int or_flags, cudaStream_t st) {
    constexpr int SMEM = WARPS * WARP_WORDS * 4;
    const unsigned grid = (unsigned)((N + WARPS - 1) / WARPS);
#define TG_ACC16_LAUNCH(KS)                                                                                            \
    {                                                                                                                  \
        auto kern = demo_acc16_mma_kernel<KS>;                                                                         \
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));                        \
        kern<<<grid, 32 * WARPS, SMEM, st>>>(tape, tape_step_stride, N, R, shift, slab, flags, or_flags);              \
    }
    switch ((R + 15) / 16) {
    case 1: TG_ACC16_LAUNCH(1) break;
    case 2: TG_ACC16_LAUNCH(2) break;
    case 3: TG_ACC16_LAUNCH(3) break;
    default: TG_ACC16_LAUNCH(4) break;
    }
#undef TG_ACC16_LAUNCH
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

} // namespace tg
