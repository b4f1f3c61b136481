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

#include "tg_demo_mma.cuh"

namespace tg {

namespace {

using namespace acc16;

constexpr int WARPS = 4;         // demos per CTA
constexpr int PREFETCH_CTAS = 5; // L2 prefetch distance in CTAs per SM (one wave) -- measured: no gain here, off by default

// KS = ceil(R / 16) K-steps.  or_flags: the flags array already holds TG_FLAG_EXHAUSTED bits of the sampler.
template <int KS, int UNR>
__global__ void __launch_bounds__(32 * WARPS)
    demo_acc16_mma_kernel(const uint8_t *__restrict__ tape, long long tape_step_stride, long long N, int R, int shift,
                          int8_t *__restrict__ slab, uint8_t *__restrict__ flags, int or_flags, int prefetch_ctas) {
    extern __shared__ __align__(128) uint32_t s_words[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n = (long long)blockIdx.x * WARPS + warp;
    if (n >= N) return;
    uint32_t *sH = s_words + warp * WARP_WORDS, *sU = sH + H_WORDS;

    // (tuning) pull the records of the demo a later wave of CTAs will sum into L2, one 48-byte record per lane and pass
    constexpr int NP = (KS + 1) / 2;
    {
        const long long np = n + (long long)prefetch_ctas * 148 * WARPS;
        if (prefetch_ctas > 0 && np < N) {
#pragma unroll
            for (int pass = 0; pass < NP; pass++)
                if (32 * pass + lane < R)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(tape + (size_t)(32 * pass + lane) * tape_step_stride + np * 48));
        }
    }
    // ---------------- 1. records -> halves: lane l takes records l and l + 32, every load in flight before the first conversion
    uint4 raw[NP][3];
#pragma unroll
    for (int pass = 0; pass < NP; pass++) {
        const int r = 32 * pass + lane;
        const uint32_t z = (uint32_t)shift * ONES4; // coefficient 0
        raw[pass][0] = raw[pass][1] = raw[pass][2] = make_uint4(z, z, z, z);
        if (r < R) {
            const uint4 *src = reinterpret_cast<const uint4 *>(tape + (size_t)r * tape_step_stride + n * 48);
#pragma unroll
            for (int q = 0; q < 3; q++) raw[pass][q] = __ldg(src + q);
        }
    }
    uint32_t invalid = 0;
#pragma unroll
    for (int pass = 0; pass < NP; pass++) {
        const int r = 32 * pass + lane;
        if (r < 16 * KS) invalid |= record_to_halves(raw[pass], shift, sH + r * HP);
    }
    int8_t *out = slab + n * 4096;
    bool bad;
    if (__any_sync(0xFFFFFFFFu, invalid != 0)) {
        // not an action list of this alphabet: entry by entry in int32, coefficient = int8(token - shift)
        uint32_t b = 0;
        for (int e = lane; e < 4096; e += 32) {
            const int i = e >> 8, j = (e >> 4) & 15, k = e & 15;
            int acc = 0;
            for (int r = 0; r < R; r++) {
                const uint8_t *rec = tape + (size_t)r * tape_step_stride + n * 48;
                acc += (int)(int8_t)(rec[i] - shift) * (int)(int8_t)(rec[16 + j] - shift) * (int)(int8_t)(rec[32 + k] - shift);
            }
            if (acc < -64 || acc > 63) b = 1;
            out[e] = (int8_t)acc;
        }
        bad = __any_sync(0xFFFFFFFFu, b != 0);
    } else {
        __syncwarp();
        bad = gemms_from_halves<KS, UNR>(sH, sU, out, lane);
    }
    if (flags && lane == 0) flags[n] = (uint8_t)((or_flags ? flags[n] : 0) | (bad ? TG_FLAG_RANGE : 0));
}

} // namespace

bool demo_acc16_mma_applies(int R) { return R >= 1 && R <= 64; }

int launch_demo_acc16_mma(const uint8_t *tape, long long tape_step_stride, long long N, int R, int shift, int8_t *slab,
                          uint8_t *flags, int or_flags, cudaStream_t st) {
    constexpr int SMEM = WARPS * WARP_WORDS * 4;
    if ((N + WARPS - 1) / WARPS > 0x7FFFFFFFLL) return TG_E_ARG;
    const unsigned grid = (unsigned)((N + WARPS - 1) / WARPS);
#ifdef TG_TUNING
    static const int variant = tuning_env("TG_ACC_VARIANT", 0);
    const int pf = (variant & 1) ? PREFETCH_CTAS : 0;
#define TG_ACC16_LAUNCH(KS)                                                                                            \
    if (variant & 2) {                                                                                                 \
        auto kern = demo_acc16_mma_kernel<KS, 2>;                                                                      \
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));                        \
        kern<<<grid, 32 * WARPS, SMEM, st>>>(tape, tape_step_stride, N, R, shift, slab, flags, or_flags, pf);          \
    } else {                                                                                                           \
        auto kern = demo_acc16_mma_kernel<KS, 4>;                                                                      \
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));                        \
        kern<<<grid, 32 * WARPS, SMEM, st>>>(tape, tape_step_stride, N, R, shift, slab, flags, or_flags, pf);          \
    }
#else
    const int pf = 0; // measured best (profiles/r01_time_demo16_mma.txt): four CTAs per SM, no L2 prefetch
#define TG_ACC16_LAUNCH(KS)                                                                                            \
    {                                                                                                                  \
        auto kern = demo_acc16_mma_kernel<KS, 4>;                                                                      \
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));                        \
        kern<<<grid, 32 * WARPS, SMEM, st>>>(tape, tape_step_stride, N, R, shift, slab, flags, or_flags, pf);          \
    }
#endif
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
