// tg_basis_mma9.cu -- K5n: change of basis of 9x9x9 games on the tensor cores
// (called by tg_change_of_basis / tg_change_of_basis_i16 for S = 9; same contract as tg_basis.cu).
//
// ABSENT from the reference; spec as in tg_basis.cu:
//     T'[i][j][k] = sum_abc A[i][a] B[j][b] C[k][c] T[a][b][c]
//
// One WARP per game, three passes of mma.sync m16n8k16 (f16 operands, f32 accumulation), the 9x9 matrix always the A
// operand (zero-padded to 16x16), the tensor the B operand.  Everything is exact integer arithmetic: an f16 holds every
// integer up to 2048 and the f32 sums stay far below 2^24; the operands of passes 2 and 3 (Y and Z) are CHECKED as they are
// produced (|.| <= 2047) and a game that fails is left to the exact int32 kernel (BASIS_REDO) -- with SURVEY 8(d)'s
// matrices (off-diagonal density 0.3) max |Z| is ~170, so that practically never happens.
//   0. the int8 game becomes rows of halves in shared memory: row n' = 9a + b holds T[a][b][0..8] (byte -> half with one
//      PRMT per two entries: 0x6400 | (byte ^ 0x80) is the half 1024 + 128 + value, one HSUB2 removes the offset);
//   1. Y[k'][a][b] = sum_c C[k'][c] T[a][b][c]: per a one ldmatrix.x4 (rows b = 0..7 and b = 8 + seven zero rows) and
//      two MMAs (N tiles b = 0..7 and b = 8..15);
//   2. Z[j'][a][k'] = sum_b B[j'][b] Y[k'][a][b]: the accumulators of the two pass-1 MMAs ARE the B fragment of pass 2
//      (column <-> k', K slots <-> b), one cvt.rn.f16x2.f32 per two values; two MMAs per a (k' = 0..7 and k' = 8);
//   3. T'[i'][j'][k'] = sum_a A[i'][a] Z[j'][a][k']: the one transposition, through shared memory: Z is stored as halves
//      [a][n = 9 j' + k'] and read back with ldmatrix.x4.trans, eleven MMAs over the 81 (j', k') columns -- whose
//      accumulators land exactly in slab order (row i', entries n, n + 1), so the result leaves as aligned pairs into a
//      staging tile and from there with one TMA bulk store (768 bytes int8, 1536 bytes int16).
#include <cuda_fp16.h>

#include "tg_common.cuh"

namespace tg {

namespace {

constexpr int WARPS = 4;             // games per CTA
constexpr int TS_PITCH = 48;         // bytes per row of halves (16 halves + pad: conflict-free ldmatrix rows)
constexpr int TS_BYTES = 82 * TS_PITCH; // 81 rows + one all-zero row = 3936
constexpr int ZS_PITCH = 208;        // bytes per row a of Z (104 halves >= 96; 52 words: conflict-free transposed reads)
constexpr int ZS_BYTES = 10 * ZS_PITCH; // rows a = 0..8 and one all-zero row (a = 9..15 all read it) = 2080
constexpr int MS_BYTES = 256;        // the game's three 9x9 int8 matrices (243 bytes at any byte alignment)
constexpr int WARP_BYTES = TS_BYTES + ZS_BYTES + MS_BYTES; // 6272
static_assert(TS_BYTES >= 1536 && WARP_BYTES % 16 == 0, "the output stage overlays the T rows");

constexpr float MAGIC = 12582912.0f; // 1.5 * 2^23: float(MAGIC + n) has n in its low mantissa bits
constexpr uint32_t MAGIC_BITS = 0x4B400000u;

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ void track_abs_max(__half2 &m, uint32_t w) { m = __hmax2(m, __habs2(*reinterpret_cast<const __half2 *>(&w))); }

// element (r, c) of matrix f (9x9 int8 at byte `base` of the staged bytes), zero outside
__device__ __forceinline__ float mat_el(const int8_t *mb, int r, int c) { return (r < 9 && c < 9) ? (float)mb[r * 9 + c] : 0.f; }

// the m16n8k16 A fragment (rows g, g + 8; columns 2t, 2t+1, 2t+8, 2t+9) of a zero-padded 9x9 matrix
__device__ __forceinline__ void mat_frag(uint32_t (&a)[4], const int8_t *mb, int g, int t) {
    a[0] = pack_f16(mat_el(mb, g, 2 * t), mat_el(mb, g, 2 * t + 1));
    a[1] = pack_f16(mat_el(mb, g + 8, 2 * t), mat_el(mb, g + 8, 2 * t + 1));
    a[2] = pack_f16(mat_el(mb, g, 2 * t + 8), mat_el(mb, g, 2 * t + 9));
    a[3] = pack_f16(mat_el(mb, g + 8, 2 * t + 8), mat_el(mb, g + 8, 2 * t + 9));
}

template <bool OUT16>
__global__ void __launch_bounds__(32 * WARPS)
    basis_mma9_kernel(const int8_t *__restrict__ slab_in, const int8_t *__restrict__ mats, long long mat_stride,
                      void *__restrict__ slab_out, uint8_t *__restrict__ flags, long long N) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const long long n = (long long)blockIdx.x * WARPS + warp;
    if (n >= N) return;
    uint8_t *s_t = smem + (size_t)warp * WARP_BYTES; // T rows, later the output stage
    uint8_t *s_z = s_t + TS_BYTES;
    uint8_t *s_m = s_z + ZS_BYTES;

    // ---- loads in flight first: the matrices (aligned words around the 243 bytes) and the game's rows
    const int8_t *mp = mats + n * mat_stride;
    const uint32_t msh = (uint32_t)(reinterpret_cast<uintptr_t>(mp) & 3);
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(mp - msh);
    const uint32_t m0 = __ldg(mw + lane), m1 = lane < 30 ? __ldg(mw + 32 + lane) : 0u; // 62 words cover 243 + 3 bytes
    const uint32_t *gw = reinterpret_cast<const uint32_t *>(slab_in + n * 768);
    uint32_t raw[3][3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const int np = lane + 32 * q; // row n' = 9a + b
        if (np < 81) {
            const int off = np * 9 + 3 * (np / 9);
#pragma unroll
            for (int w = 0; w < 3; w++) raw[q][w] = __ldg(gw + (off >> 2) + w);
        } else {
            raw[q][0] = raw[q][1] = raw[q][2] = 0;
        }
    }
    // zero row of T, zero row of Z for a = 9..15 (their A-operand columns are zero, but 0 * NaN is not)
    if (lane < 3) reinterpret_cast<uint4 *>(s_t + 81 * TS_PITCH)[lane] = make_uint4(0, 0, 0, 0);
    if (lane < ZS_PITCH / 16) reinterpret_cast<uint4 *>(s_z + 9 * ZS_PITCH)[lane] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint32_t *>(s_m)[lane] = m0;
    reinterpret_cast<uint32_t *>(s_m)[32 + lane] = m1;

    // ---- 0. T[a][b][.] as rows of halves
    constexpr uint32_t KB = 0x00006480u; // bytes 0x80, 0x64 for PRMT: 0x64xx is the half 1024 + xx, 0x6480 the half 1152
    const __half2 ks = __halves2half2(__ushort_as_half((unsigned short)0x6480), __ushort_as_half((unsigned short)0x6480));
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const int np = lane + 32 * q;
        if (np < 81) {
            const int off = np * 9 + 3 * (np / 9);
            const uint32_t sh = 8 * (off & 3);
            const uint32_t x0 = __funnelshift_r(raw[q][0], raw[q][1], sh) ^ H4, x1 = __funnelshift_r(raw[q][1], raw[q][2], sh) ^ H4,
                           x2 = (raw[q][2] >> sh) ^ H4;
            uint32_t h[5] = {prmt(x0, KB, 0x5150u), prmt(x0, KB, 0x5352u), prmt(x1, KB, 0x5150u), prmt(x1, KB, 0x5352u), prmt(x2, KB, 0x5450u)};
#pragma unroll
            for (int e = 0; e < 5; e++) {
                const __half2 v = __hsub2(*reinterpret_cast<const __half2 *>(&h[e]), ks);
                h[e] = *reinterpret_cast<const uint32_t *>(&v);
            }
            uint4 *row = reinterpret_cast<uint4 *>(s_t + np * TS_PITCH);
            row[0] = make_uint4(h[0], h[1], h[2], h[3]);
            row[1] = make_uint4(h[4], 0, 0, 0);
        }
    }
    __syncwarp();
    uint32_t fA[4], fB[4], fC[4];
    {
        const int8_t *mb = reinterpret_cast<const int8_t *>(s_m) + msh;
        mat_frag(fA, mb, g, t);
        mat_frag(fB, mb + 81, g, t);
        mat_frag(fC, mb + 162, g, t);
    }

    // ---- 1 + 2, one a at a time
    __half2 ymax = __float2half2_rn(0.f), zmax = __float2half2_rn(0.f);
    const uint32_t ts_base = smem_u32(s_t), zs_base = smem_u32(s_z);
    // ldmatrix row addresses of this lane: matrix = lane >> 3 (bit 0: halves 8..15, bit 1: rows b = 8..15), row = lane & 7
    const int lm_b = (lane & 7) + 8 * (lane >> 4), lm_c = ((lane >> 3) & 1) * 16;
    const bool lm_real = lm_b < 9;
    uint16_t *zrow = reinterpret_cast<uint16_t *>(s_z);
#pragma unroll
    for (int a = 0; a < 9; a++) {
        uint32_t tb[4];
        ldmatrix_x4(tb, ts_base + (lm_real ? (9 * a + lm_b) : 81) * TS_PITCH + lm_c);
        float y0[4] = {0.f, 0.f, 0.f, 0.f}, y1[4] = {0.f, 0.f, 0.f, 0.f}; // rows k' = g | g+8 ; cols b = 2t, 2t+1 | 8+2t, 9+2t
        mma_f16(y0, fC, tb[0], tb[1]);
        mma_f16(y1, fC, tb[2], tb[3]);
        const uint32_t b00 = pack_f16(y0[0], y0[1]), b01 = pack_f16(y1[0], y1[1]); // k' = g
        const uint32_t b10 = pack_f16(y0[2], y0[3]), b11 = pack_f16(y1[2], y1[3]); // k' = g + 8
        track_abs_max(ymax, b00), track_abs_max(ymax, b01), track_abs_max(ymax, b10), track_abs_max(ymax, b11);
        float z0[4] = {0.f, 0.f, 0.f, 0.f}, z1[4] = {0.f, 0.f, 0.f, 0.f}; // rows j' = g | g+8 ; cols k' = 2t, 2t+1 | 8+2t, 9+2t
        mma_f16(z0, fB, b00, b01);
        mma_f16(z1, fB, b10, b11);
        // Z as halves [a][n = 9 j' + k']
        const uint32_t p0 = pack_f16(z0[0], z0[1]), p1 = pack_f16(z0[2], z0[3]), p2 = pack_f16(z1[0], z1[2]);
        track_abs_max(zmax, p0), track_abs_max(zmax, p1), track_abs_max(zmax, p2);
        uint16_t *zr = zrow + a * (ZS_PITCH / 2);
        zr[9 * g + 2 * t] = (uint16_t)p0, zr[9 * g + 2 * t + 1] = (uint16_t)(p0 >> 16); // j' = g, k' = 2t, 2t+1
        if (g == 0) *reinterpret_cast<uint32_t *>(zr + 72 + 2 * t) = p1;               // j' = 8
        if (t == 0) zr[9 * g + 8] = (uint16_t)p2;                                       // j' = g, k' = 8
        if (lane == 0) zr[80] = (uint16_t)(p2 >> 16);                                   // j' = 8, k' = 8
    }
    {
        const float ym = fmaxf(__low2float(ymax), __high2float(ymax)), zm = fmaxf(__low2float(zmax), __high2float(zmax));
        // 2047 is the largest bound below which every integer is an f16 (a value that rounded is >= 2048 after rounding too)
        if (__any_sync(0xFFFFFFFFu, !(ym <= 2047.f && zm <= 2047.f))) {
            if (lane == 0) flags[n] = BASIS_REDO;
            return;
        }
    }
    __syncwarp(); // Z complete; the T rows are dead: their bytes become the output stage

    // ---- 3. eleven N tiles over n = 9 j' + k'; accumulator (row i', cols n, n+1) is slab entry i' * 84 + n
    uint32_t over = 0;
    const int zl_row = min((lane & 7) + 8 * ((lane >> 3) & 1), 9), zl_col = (lane >> 4) * 16; // ldmatrix.trans: rows a (9 = the zero row), 16 bytes of columns
#pragma unroll
    for (int tp = 0; tp < 6; tp++) {
        uint32_t zb[4];
        ldmatrix_x4_trans(zb, zs_base + zl_row * ZS_PITCH + tp * 32 + zl_col);
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            const int tile = 2 * tp + hh;
            if (tile >= 11) break;
            constexpr float M0 = OUT16 ? MAGIC + 32768.f : MAGIC;
            float d[4] = {M0, M0, M0, M0};
            mma_f16(d, fA, zb[2 * hh], zb[2 * hh + 1]);
            const int nn = 8 * tile + 2 * t; // entries nn, nn + 1 of row i' = g (d[0], d[1]) and i' = g + 8 (d[2], d[3])
            // the last tile holds entry 80 and, beyond it, the row padding (81..83), which is stored as zero
            const bool last = tile == 10;
            const bool v0 = !last || t == 0, v1 = !last, st = !last || t < 2;
            const uint32_t u0 = __float_as_uint(d[0]), u1 = __float_as_uint(d[1]), u2 = __float_as_uint(d[2]), u3 = __float_as_uint(d[3]);
            if constexpr (OUT16) {
                // low half = result in offset binary, high half = 0x4B40 iff it fits int16
                const uint32_t keep = last ? (v0 ? 0xFFFFu : 0u) : 0xFFFFFFFFu;
                const uint32_t wA = (prmt(u0, u1, 0x5410u) ^ 0x80008000u) & keep, wB = (prmt(u2, u3, 0x5410u) ^ 0x80008000u) & keep;
                over |= (prmt(u0, u1, 0x7632u) ^ 0x4B404B40u) & keep;
                if (g == 0) over |= (prmt(u2, u3, 0x7632u) ^ 0x4B404B40u) & keep;
                if (st) {
                    reinterpret_cast<uint32_t *>(s_t)[(g * 84 + nn) >> 1] = wA;
                    if (g == 0) reinterpret_cast<uint32_t *>(s_t)[(8 * 84 + nn) >> 1] = wB;
                }
            } else {
                // in [-64, 63]  <=>  bits - (MAGIC_BITS - 64) < 128
                uint32_t o = 0;
                if (v0) o |= (u0 - (MAGIC_BITS - 64u)) & ~127u;
                if (v1) o |= (u1 - (MAGIC_BITS - 64u)) & ~127u;
                if (g == 0 && v0) o |= (u2 - (MAGIC_BITS - 64u)) & ~127u;
                if (g == 0 && v1) o |= (u3 - (MAGIC_BITS - 64u)) & ~127u;
                over |= o;
                const uint32_t keep = last ? (v0 ? 0xFFu : 0u) : 0xFFFFu;
                if (st) {
                    reinterpret_cast<uint16_t *>(s_t)[(g * 84 + nn) >> 1] = (uint16_t)(prmt(u0, u1, 0x4040u) & keep);
                    if (g == 0) reinterpret_cast<uint16_t *>(s_t)[(8 * 84 + nn) >> 1] = (uint16_t)(prmt(u2, u3, 0x4040u) & keep);
                }
            }
        }
    }
    // game padding (entries 756..767) is zero
    if (lane < (OUT16 ? 6 : 3)) reinterpret_cast<uint32_t *>(s_t)[(OUT16 ? 378 : 189) + lane] = 0u;
    const bool bad = __any_sync(0xFFFFFFFFu, over != 0);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        flags[n] = bad ? (uint8_t)TG_FLAG_RANGE : (uint8_t)0;
        constexpr uint32_t OB = OUT16 ? 1536u : 768u;
        bulk_s2g(reinterpret_cast<uint8_t *>(slab_out) + n * OB, s_t, OB);
        bulk_commit();
        bulk_wait_read<0>();
    }
}

} // namespace

int launch_basis_mma9(const int8_t *slab_in, const int8_t *mats, long long mat_stride, void *slab_out, int out16, uint8_t *flags,
                      long long N, cudaStream_t st) {
    constexpr int SMEM = WARPS * WARP_BYTES;
    const unsigned grid = (unsigned)((N + WARPS - 1) / WARPS);
    if (out16) {
        auto kern = basis_mma9_kernel<true>;
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        kern<<<grid, 32 * WARPS, SMEM, st>>>(slab_in, mats, mat_stride, slab_out, flags, N);
    } else {
        auto kern = basis_mma9_kernel<false>;
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        kern<<<grid, 32 * WARPS, SMEM, st>>>(slab_in, mats, mat_stride, slab_out, flags, N);
    }
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

} // namespace tg
