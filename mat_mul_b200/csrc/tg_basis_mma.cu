// tg_basis_mma.cu -- K5t: change of basis of 16x16x16 games on the tensor cores
// (called by tg_change_of_basis for S = 16; same contract as tg_basis.cu).
//
// ABSENT from the reference; spec as in tg_basis.cu:
//     T'[i][j][k] = sum_abc A[i][a] B[j][b] C[k][c] T[a][b][c]
//
// One WARP per game, three passes of mma.sync m16n8k16, the 16x16 matrix always
// the A operand (M = the new index), the tensor the B operand (K = the contracted
// index, N = 8 values of one surviving index):
//   1. Y[k'][a][b] = sum_c C[k'][c] T[a][b][c]   int8 MMA; B fragment = a slab word as it lies in
//      HBM (lane (g,t) reads word (a*16 + 8h + g)*4 + t: one coalesced 128-byte line per MMA);
//   2. Z[j'][a][k'] = sum_b B[j'][b] Y[k'][a][b]  the accumulators of two pass-1 MMAs ARE the
//      next B fragment (column g <-> k' = g; K slots <-> the b of the two tiles);
//   3. T'[i'][j'][k'] = sum_a A[i'][a] Z[j'][a][k']  needs a on the K slots of a lane while
//      pass 2 leaves it spread over MMAs: the one real transposition, through 11 KB of
//      shared memory per warp (16 STS.128 + 16 LDS.128 per lane, conflict-free pitches).
// Y and Z do not fit int8.  Two ways to feed them to the next pass, chosen per game
// (warp-uniform), both bit-identical to the int64 oracle:
//   F16: passes 2 and 3 are f16 MMAs with f32 accumulation -- exact integer arithmetic while
//        |Y|, |Z| <= 2048 (f16 holds those integers, the f32 sums stay below 2^24).  The int32
//        accumulators of pass 1 start at the bit pattern of 1.5 * 2^23, so they are floats already
//        (one FADD removes the offset); accumulators become the next operand with one
//        cvt.rn.f16x2.f32 per two values; pass 3 starts at 1.5 * 2^23 again, so the low byte of its
//        accumulators is the int8 result.  Guard (a priori, tb >= max|T| from the OR of the byte
//        magnitudes):  tb ||C|| <= 2048,  tb ||C|| ||B|| <= 2048,  tb ||C|| ||B|| ||A|| <= 32767.
//   P16: int8 MMAs on the low 16 bits as two byte planes (arithmetic modulo 2^16 is a ring
//        homomorphism):  M x = M lo8(x) [s8 x u8] + 256 * M hi8(x) [s8 x s8], two MMAs chained through
//        the accumulator.  T' mod 256 is always right; the range flag reads the 16-bit results and
//        is exact iff |T'| <= 32767, guaranteed by ||C|| <= 255 (Y fits 16 bits) and
//        ybound ||A|| ||B|| <= 32767 with ybound >= max|Y| read off the high byte plane.
//   else the game is marked BASIS_REDO and redone by the exact int32 kernel of tg_basis.cu.
// Why not int8 MMAs throughout: every legacy MMA shape issues at 0.49 per clock and SM
// (scripts/ubench_pipes.cu), and the byte-plane marshaling (PRMT) saturates the ALU pipe;
// cvt.rn.f16x2.f32 issues beside it (profiles/README.md).
#include <cuda_fp16.h>

#include "tg_common.cuh"

namespace tg {

namespace {

// transposition buffer of a warp: word (j', slot, pos), slot = plane * 4 + a / 4 (P16) or a / 2 (F16), pos = 4t + the
// lane's k' position; pitches chosen so that neither the 128-bit stores nor the 128-bit loads conflict
constexpr int PQ = 20;              // words between slots: 16 + 4
constexpr int PJ = 8 * PQ + 16;     // words between j' rows (176 = 16 mod 32)
constexpr int WARP_WORDS = 16 * PJ; // 2816 words = 11 KB per warp
constexpr int WARPS = 4;            // games per CTA

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// d = A(s8) * B(s8) + c
__device__ __forceinline__ void mma_ss(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
// d = A(s8) * B(u8) + c
__device__ __forceinline__ void mma_su(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}

// M * x for x given as byte planes (lo unsigned, hi signed), modulo 2^16 in the low half of d
__device__ __forceinline__ void mma_planes(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t lo, uint32_t hi) {
    d[0] = d[1] = d[2] = d[3] = 0;
    mma_ss(d, a0, a1, hi);
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] *= 256;
    mma_su(d, a0, a1, lo);
}

// four int32 -> their low bytes in one word (lo) and their second bytes in another (hi)
__device__ __forceinline__ void pack_planes(int v0, int v1, int v2, int v3, uint32_t &lo, uint32_t &hi) {
    const uint32_t p01 = prmt((uint32_t)v0, (uint32_t)v1, 0x5140u), p23 = prmt((uint32_t)v2, (uint32_t)v3, 0x5140u);
    lo = prmt(p01, p23, 0x5410u);
    hi = prmt(p01, p23, 0x7632u);
}

// d += A(f16, 16x16) * B(f16, 16x8), f32 accumulation
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (lo, hi) -> f16x2 word, lo in the low half
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// int8 bytes b0, b1 of w (as selected by the caller) -> f16x2
__device__ __forceinline__ uint32_t bytes_to_f16(uint32_t w, int b0, int b1) {
    return pack_f16((float)(int8_t)(w >> (8 * b0)), (float)(int8_t)(w >> (8 * b1)));
}

constexpr float MAGIC = 12582912.0f;     // 1.5 * 2^23: float(MAGIC + n) has n in its low mantissa bits
constexpr int MAGIC_BITS = 0x4B400000;

// one's complement magnitude of the four int8 of w (|x| <= mag + 1), for OR-accumulated bounds
__device__ __forceinline__ uint32_t mag4(uint32_t w) { return w ^ prmt(w, 0u, 0xBA98u); }

// OR of the four bytes of w, over the warp
__device__ __forceinline__ int warp_or_bytes(uint32_t w) {
    w |= w >> 16;
    w |= w >> 8;
    return (int)__reduce_or_sync(0xFFFFFFFFu, w & 0xFFu);
}

// sum of |byte| over the four int8 of w
__device__ __forceinline__ int abs_sum4(uint32_t w) { return __dp4a((int)w, (int)(prmt(w, 0u, 0xBA98u) | ONES4), 0); }

// ||M||inf of a matrix whose fragment (rows g and g+8, four columns per lane) is (a0, a1)
__device__ __forceinline__ int norm_inf(uint32_t a0, uint32_t a1) {
    int r0 = abs_sum4(a0), r1 = abs_sum4(a1);
    r0 += __shfl_xor_sync(0xFFFFFFFFu, r0, 1), r1 += __shfl_xor_sync(0xFFFFFFFFu, r1, 1);
    r0 += __shfl_xor_sync(0xFFFFFFFFu, r0, 2), r1 += __shfl_xor_sync(0xFFFFFFFFu, r1, 2);
    return (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)max(r0, r1));
}

// range test + int8 packing of sixteen results held as the low 16 bits of d[k' & 3][x]: in [-64, 63] <=> bits 6..15 all equal
__device__ __forceinline__ void finish4(const uint32_t (&d)[4][4], uint32_t &over, uint32_t (&outw)[4][4], int kq) {
#pragma unroll
    for (int x = 0; x < 4; x++) {
        const uint32_t w01 = prmt(d[0][x], d[1][x], 0x5410u), w23 = prmt(d[2][x], d[3][x], 0x5410u);
        over |= ((w01 ^ (w01 + w01)) | (w23 ^ (w23 + w23))) & 0xFF80FF80u;
        outw[x][kq] = prmt(w01, w23, 0x6420u);
    }
}

// int16 out: sixteen results d[k' & 3][x] -> eight int16 pairs.  F16: the accumulators started at 1.5 * 2^23 + 2^15, so the low
// half of the bit pattern is the result in offset binary and the high half is 0x4B40 iff the result fits int16; P16: the low
// half is the result modulo 2^16 (its guard already proved that it fits).
template <bool F16>
__device__ __forceinline__ void finish4_i16(const uint32_t (&d)[4][4], uint32_t &over, uint32_t (&outw)[4][8], int kq) {
#pragma unroll
    for (int x = 0; x < 4; x++) {
        const uint32_t w01 = prmt(d[0][x], d[1][x], 0x5410u), w23 = prmt(d[2][x], d[3][x], 0x5410u);
        if constexpr (F16) {
            over |= (prmt(d[0][x], d[1][x], 0x7632u) ^ 0x4B404B40u) | (prmt(d[2][x], d[3][x], 0x7632u) ^ 0x4B404B40u);
            outw[x][2 * kq] = w01 ^ 0x80008000u, outw[x][2 * kq + 1] = w23 ^ 0x80008000u;
        } else {
            outw[x][2 * kq] = w01, outw[x][2 * kq + 1] = w23;
        }
    }
}

// |h| of both halves of an f16x2 word, folded into a running maximum (exactness check of the f16 operands)
__device__ __forceinline__ void track_abs_max(__half2 &m, uint32_t w) {
    m = __hmax2(m, __habs2(*reinterpret_cast<const __half2 *>(&w)));
}

// everything after the loads, for one game held by one warp.  fm: fragments of C (int8), of A and B with the columns
// permuted to 2t, 2t+1, 8+2t, 9+2t (int8; the K-slot order of P16's pass 2 and the source of the f16 fragments) and of A
// with its natural columns 4t..4t+3 (P16's pass 3).
struct Frags {
    uint32_t c0, c1, pa0, pa1, pb0, pb1, na0, na1;
};

// OUT16: int16 slab out.  CHECK (F16 only): the a-priori norm guard did not hold, so every f16 operand is checked as it is
// produced (|Y|, |Z| <= 2047 keeps the f16 conversion exact, and then every f32 sum is exact too); returns false -- before
// anything has been written -- if one is not, and the caller falls back to the byte-plane path.
template <bool F16, bool STREAM, bool OUT16, bool CHECK>
__device__ __forceinline__ bool basis_mma16_game(uint32_t (&tw)[32], const uint32_t *__restrict__ src, const Frags &fm, int nA,
                                                 int nB, int nC, uint32_t *sw, void *__restrict__ out_v,
                                                 uint8_t *__restrict__ flag, int lane) {
    const int g = lane >> 2, t = lane & 3;
    __half2 opmax = __float2half2_rn(0.f), zmax = __float2half2_rn(0.f); // CHECK: largest |Y| and |Z| operand of this lane
    uint32_t hA[4], hB[4]; // F16: the f16 fragments (rows g | g+8, k = 2t, 2t+1 | 2t+8, 2t+9)
    if constexpr (F16) {
        hA[0] = bytes_to_f16(fm.pa0, 0, 1), hA[1] = bytes_to_f16(fm.pa1, 0, 1);
        hA[2] = bytes_to_f16(fm.pa0, 2, 3), hA[3] = bytes_to_f16(fm.pa1, 2, 3);
        hB[0] = bytes_to_f16(fm.pb0, 0, 1), hB[1] = bytes_to_f16(fm.pb1, 0, 1);
        hB[2] = bytes_to_f16(fm.pb0, 2, 3), hB[3] = bytes_to_f16(fm.pb1, 2, 3);
    }
    // ---------------- passes 1 and 2, four a at a time; Z leaves packed along a
    uint32_t ymag = 0; // P16: OR of the magnitudes of the high bytes of Y
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if constexpr (STREAM) {
            if (q == 1 || q == 2) {
#pragma unroll
                for (int x = 0; x < 8; x++) tw[8 * (q + 1) + x] = __ldg(src + (8 * (q + 1) + x) * 32 + lane);
            }
        }
        int zi[4][2][4];   // P16  [a & 3][k' half][2 * (j' half) + (k' & 1)]
        float zf[4][2][4]; // F16
#pragma unroll
        for (int aa = 0; aa < 4; aa++) {
            const int a = 4 * q + aa;
            constexpr int C0 = F16 ? MAGIC_BITS : 0;
            int y0[4] = {C0, C0, C0, C0}, y1[4] = {C0, C0, C0, C0}; // b = 2t, 2t+1 | 8+2t, 9+2t ; rows k' = g | g+8
            mma_ss(y0, fm.c0, fm.c1, tw[2 * a]);
            mma_ss(y1, fm.c0, fm.c1, tw[2 * a + 1]);
#pragma unroll
            for (int hk = 0; hk < 2; hk++) {
                if constexpr (F16) {
                    const uint32_t b0 = pack_f16(__int_as_float(y0[2 * hk]) - MAGIC, __int_as_float(y0[2 * hk + 1]) - MAGIC);
                    const uint32_t b1 = pack_f16(__int_as_float(y1[2 * hk]) - MAGIC, __int_as_float(y1[2 * hk + 1]) - MAGIC);
                    if constexpr (CHECK) track_abs_max(opmax, b0), track_abs_max(opmax, b1);
                    zf[aa][hk][0] = zf[aa][hk][1] = zf[aa][hk][2] = zf[aa][hk][3] = 0.f;
                    mma_f16(zf[aa][hk], hB, b0, b1);
                } else {
                    uint32_t lo, hi;
                    pack_planes(y0[2 * hk], y0[2 * hk + 1], y1[2 * hk], y1[2 * hk + 1], lo, hi);
                    ymag |= mag4(hi);
                    mma_planes(zi[aa][hk], fm.pb0, fm.pb1, lo, hi);
                }
            }
        }
#pragma unroll
        for (int hj = 0; hj < 2; hj++) {
            uint32_t w0[4], w1[4]; // position 2 * (k' half) + (k' & 1)  <->  k' = 8 hk + 2t + ek
#pragma unroll
            for (int hk = 0; hk < 2; hk++)
#pragma unroll
                for (int ek = 0; ek < 2; ek++) {
                    const int x = 2 * hj + ek;
                    if constexpr (F16) { // slots a / 2 = 2q, 2q + 1
                        w0[2 * hk + ek] = pack_f16(zf[0][hk][x], zf[1][hk][x]);
                        w1[2 * hk + ek] = pack_f16(zf[2][hk][x], zf[3][hk][x]);
                        if constexpr (CHECK) track_abs_max(zmax, w0[2 * hk + ek]), track_abs_max(zmax, w1[2 * hk + ek]);
                    } else { // slots q (low bytes), 4 + q (high bytes)
                        pack_planes(zi[0][hk][x], zi[1][hk][x], zi[2][hk][x], zi[3][hk][x], w0[2 * hk + ek], w1[2 * hk + ek]);
                    }
                }
            uint32_t *dst = sw + (g + 8 * hj) * PJ + 4 * t;
            *reinterpret_cast<uint4 *>(dst + (F16 ? 2 * q : q) * PQ) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
            *reinterpret_cast<uint4 *>(dst + (F16 ? 2 * q + 1 : 4 + q) * PQ) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        }
    }
    if constexpr (!F16) {
        // exactness guard (see the header): every lane computes the same verdict
        const long long ybound = 256LL * ((long long)warp_or_bytes(ymag) + 1);
        if (!(nC <= 255 && ybound * nA * nB <= 32767)) {
            if (lane == 0) *flag = BASIS_REDO;
            return true;
        }
    }
    if constexpr (F16 && CHECK) {
        // 2047 is the largest integer below which every integer is an f16; a value that rounded is >= 2048 after rounding too
        const float ym = fmaxf(__low2float(opmax), __high2float(opmax)), zm = fmaxf(__low2float(zmax), __high2float(zmax));
        bool viol = !(ym <= 2047.f && zm <= 2047.f);
        // int8 out: the range test below reads the low 16 bits of the results, so they must fit 16 bits: |T'| <= max|Z| ||A||
        if constexpr (!OUT16) viol |= !(zm * (float)nA <= 32767.f);
        if (__any_sync(0xFFFFFFFFu, viol)) return false;
    }
    __syncwarp();

    // ---------------- pass 3: lane (g,t) now supplies the a of its K slots for column j' = g (+8) and every k'
    uint32_t over = 0;
#pragma unroll
    for (int hj = 0; hj < 2; hj++) {
        uint32_t r0[16], r1[16]; // slots t and t + 4; index 4s + 2hk + ek  <->  k' = 8 hk + 2s + ek
        const uint32_t *rsrc = sw + (g + 8 * hj) * PJ + t * PQ;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const uint4 l4 = *reinterpret_cast<const uint4 *>(rsrc + 4 * s), h4 = *reinterpret_cast<const uint4 *>(rsrc + 4 * PQ + 4 * s);
            r0[4 * s] = l4.x, r0[4 * s + 1] = l4.y, r0[4 * s + 2] = l4.z, r0[4 * s + 3] = l4.w;
            r1[4 * s] = h4.x, r1[4 * s + 1] = h4.y, r1[4 * s + 2] = h4.z, r1[4 * s + 3] = h4.w;
        }
        uint32_t outw[4][OUT16 ? 8 : 4]; // [2 * (i' half) + (j' & 1)][k' / 4 (int8) or k' / 2 (int16)]
#pragma unroll
        for (int kq = 0; kq < 4; kq++) {
            uint32_t d[4][4]; // [k' & 3][2 * (i' half) + (j' & 1)]
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int kp = 4 * kq + e, idx = 4 * ((kp & 7) >> 1) + 2 * (kp >> 3) + (kp & 1);
                if constexpr (F16) {
                    constexpr float M0 = OUT16 ? MAGIC + 32768.f : MAGIC;
                    float f[4] = {M0, M0, M0, M0};
                    mma_f16(f, hA, r0[idx], r1[idx]);
#pragma unroll
                    for (int x = 0; x < 4; x++) d[e][x] = __float_as_uint(f[x]);
                } else {
                    int v[4];
                    mma_planes(v, fm.na0, fm.na1, r0[idx], r1[idx]);
#pragma unroll
                    for (int x = 0; x < 4; x++) d[e][x] = (uint32_t)v[x];
                }
            }
            if constexpr (OUT16)
                finish4_i16<F16>(d, over, outw, kq);
            else
                finish4(d, over, outw, kq);
        }
#pragma unroll
        for (int x = 0; x < 4; x++) {
            const int ip = g + 8 * (x >> 1), jp = 8 * hj + 2 * t + (x & 1);
            if constexpr (OUT16) {
                uint4 *o = reinterpret_cast<uint4 *>(reinterpret_cast<int16_t *>(out_v) + ip * 256 + jp * 16);
                o[0] = make_uint4(outw[x][0], outw[x][1], outw[x][2], outw[x][3]);
                o[1] = make_uint4(outw[x][4], outw[x][5], outw[x][6], outw[x][7]);
            } else {
                *reinterpret_cast<uint4 *>(reinterpret_cast<int8_t *>(out_v) + ip * 256 + jp * 16) =
                    make_uint4(outw[x][0], outw[x][1], outw[x][2], outw[x][3]);
            }
        }
    }
    const bool bad = __any_sync(0xFFFFFFFFu, over != 0);
    if (lane == 0) *flag = (uint8_t)((bad ? TG_FLAG_RANGE : 0u) | (F16 ? 0u : TG_FLAG_PATH_PLANES));
    return true;
}

// MINB: CTAs per SM the register budget is sized for; STREAM: load the slab words four a at a time (one group ahead)
// instead of all 32 up front; PATHS: 0 = F16 where its guard holds, else P16 (default), 1 = P16 only (tuning);
// PREFETCH: L2 prefetch distance in CTAs per SM (0 = none)
template <int MINB, bool STREAM, int PATHS, int PREFETCH, bool OUT16>
__global__ void __launch_bounds__(32 * WARPS, MINB)
    basis_mma16_kernel(const int8_t *__restrict__ slab_in, const int8_t *__restrict__ mats, long long mat_stride,
                       void *__restrict__ slab_out, uint8_t *__restrict__ flags, long long N) {
    extern __shared__ __align__(16) uint32_t s_words[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const long long n = (long long)blockIdx.x * WARPS + warp;
    if (n >= N) return;
    uint32_t *sw = s_words + warp * WARP_WORDS;

    // the game: coalesced word loads, in flight before the first MMA
    const uint32_t *src = reinterpret_cast<const uint32_t *>(slab_in + n * 4096);
    uint32_t tw[32];
#pragma unroll
    for (int q = 0; q < (STREAM && PATHS != 0 ? 16 : 32); q++) tw[q] = __ldg(src + q * 32 + lane);
    // pull the game a later wave of CTAs will work on (and its matrices) into L2: one line per lane
    if constexpr (PREFETCH > 0) {
        const long long np = n + (long long)PREFETCH * 148 * WARPS;
        if (np < N) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(slab_in + np * 4096 + lane * 128));
            if (lane < 6 && mat_stride) asm volatile("prefetch.global.L2 [%0];" ::"l"(mats + np * mat_stride + lane * 128));
        }
    }
    // matrix fragments: rows g and g+8.  Where a pass consumes accumulators directly, the K slots of a lane hold the
    // contracted indices 2t, 2t+1, 8+2t, 9+2t, so the matrix columns are gathered in that order
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(mats + n * mat_stride);
    const uint32_t sel = (t & 1) ? 0x7632u : 0x5410u;
    Frags fm;
    fm.c0 = __ldg(mw + 128 + g * 4 + t), fm.c1 = __ldg(mw + 128 + (g + 8) * 4 + t);
    fm.na0 = __ldg(mw + g * 4 + t), fm.na1 = __ldg(mw + (g + 8) * 4 + t);
    fm.pa0 = prmt(__ldg(mw + g * 4 + (t >> 1)), __ldg(mw + g * 4 + 2 + (t >> 1)), sel);
    fm.pa1 = prmt(__ldg(mw + (g + 8) * 4 + (t >> 1)), __ldg(mw + (g + 8) * 4 + 2 + (t >> 1)), sel);
    fm.pb0 = prmt(__ldg(mw + 64 + g * 4 + (t >> 1)), __ldg(mw + 64 + g * 4 + 2 + (t >> 1)), sel);
    fm.pb1 = prmt(__ldg(mw + 64 + (g + 8) * 4 + (t >> 1)), __ldg(mw + 64 + (g + 8) * 4 + 2 + (t >> 1)), sel);
    const int nA = norm_inf(fm.pa0, fm.pa1), nB = norm_inf(fm.pb0, fm.pb1), nC = norm_inf(fm.c0, fm.c1);
    void *out = reinterpret_cast<uint8_t *>(slab_out) + n * 4096 * (OUT16 ? 2 : 1);
    if constexpr (PATHS == 0) {
        // the a-priori bound needs max|T|: the whole game is read first
        uint32_t tm = 0;
#pragma unroll
        for (int q = 0; q < 32; q++) tm |= mag4(tw[q]);
        const long long yb = ((long long)warp_or_bytes(tm) + 1) * nC, zb = yb * nB;
        if (yb <= 2047 && zb <= 2047 && (OUT16 || zb * nA <= 32767)) { // every operand is an exact f16 whatever the game
            basis_mma16_game<true, false, OUT16, false>(tw, src, fm, nA, nB, nC, sw, out, flags + n, lane);
            return;
        }
        // bigger matrices (SURVEY 8(d)'s density 0.3): the bound is far from tight, so try f16 and check the operands
        if (basis_mma16_game<true, false, OUT16, true>(tw, src, fm, nA, nB, nC, sw, out, flags + n, lane)) return;
        basis_mma16_game<false, false, OUT16, false>(tw, src, fm, nA, nB, nC, sw, out, flags + n, lane);
    } else {
        basis_mma16_game<false, STREAM, OUT16, false>(tw, src, fm, nA, nB, nC, sw, out, flags + n, lane);
    }
}

} // namespace

int launch_basis_mma16(const int8_t *slab_in, const int8_t *mats, long long mat_stride, void *slab_out, int out16, uint8_t *flags,
                       long long N, cudaStream_t st) {
    constexpr int SMEM = WARPS * WARP_WORDS * 4;
    const unsigned grid = (unsigned)((N + WARPS - 1) / WARPS);
#define TG_MMA16_LAUNCH(MINB, STREAM, PATHS, PF)                                                                        \
    {                                                                                                                  \
        auto kern = out16 ? basis_mma16_kernel<MINB, STREAM, PATHS, PF, true> : basis_mma16_kernel<MINB, STREAM, PATHS, PF, false>; \
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));                        \
        kern<<<grid, 32 * WARPS, SMEM, st>>>(slab_in, mats, mat_stride, slab_out, flags, N);                           \
    }
#ifdef TG_TUNING
    switch (tuning_env("TG_BASIS_VARIANT", 0)) {
    case 2: TG_MMA16_LAUNCH(4, true, 1, 0) return TG_OK;  // 16-bit planes only
    case 3: TG_MMA16_LAUNCH(4, false, 0, 0) return TG_OK; // no L2 prefetch
    default: break;
    }
#endif
    TG_MMA16_LAUNCH(4, false, 0, 4) // measured best (profiles/README.md): prefetch one wave of CTAs ahead
#undef TG_MMA16_LAUNCH
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

} // namespace tg
