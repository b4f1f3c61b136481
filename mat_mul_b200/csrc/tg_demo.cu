// tg_demo.cu -- K3: synthetic demonstrations (C ABI: tg_demo_gen_philox,
// tg_demo_accumulate, tg_demo_from_ustream, tg_mt19937_fill_f64).
//
// Reference restated: create_synthetic_demo (utils.py:203-233) ==
// SyntheticDemoDataset._create_synthetic_demos (datasets.py:124-142): R
// accepted random factor triples (rejected iff u(x)v(x)w == 0), tokens =
// factors + shift, target = sum of the R rank-1 tensors.
//
// Device formats: multi-step tape uint8 [R][N][TP] (step-major, so step r of
// all games is one contiguous tg_step operand) and the int8 slab [N][GP].
//
// One CTA builds a tile of TG demos:
//   A. (throughput mode) all threads draw factor triples with Philox4x32-10,
//      one (demo, term) pair at a time from a shared work counter, retrying
//      rejected triples, and write the tokens into a shared-memory tape;
//      (replay modes) the tape tile is bulk-loaded from HBM instead.
//   B. WR threads per demo accumulate the R rank-1 terms of their word column
//      in registers: acc_i += u_{r,i} * pack(v_r w_r) -- one IMAD per four
//      entries per term (same packed arithmetic as tg_step.cuh), with a range
//      check often enough that the packed form can never alias.
//   C. the tile (slab + tape) leaves through TMA bulk stores.
#include "tg_step.cuh"

namespace tg {

struct Categorical {
    uint32_t thr[8]; // 16-bit CDF thresholds (65536 = never exceeded)
    int8_t values[8];
    int n;
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

template <int S, int NT, int NPASS>
struct DemoCfg {
    using G = Geo<S>;
    static constexpr int GPASS = NT / G::WR;
    static constexpr int TG = GPASS * NPASS;
    static constexpr int ACTIVE = GPASS * G::WR;
    static constexpr int SLAB_BYTES = TG * G::GP;
    static constexpr int PW = (G::WR % 32 == 0) ? G::WR / 32 : ((32 % G::WR == 0) ? 1 : G::WR);
    static __host__ __device__ constexpr int tape_bytes(int R) { return R * TG * G::TP; }
    static __host__ __device__ constexpr int smem_bytes(int R) {
        return SLAB_BYTES + tape_bytes(R) + TG * PW * 4 + TG * 4 + 16;
    }
};

// draw one factor triple (3S tokens) for (demo d, term r, try t); returns true if accepted
template <int S>
__device__ __forceinline__ bool draw_triple(uint32_t words[Geo<S>::TP / 4], uint32_t k0, uint32_t k1, uint32_t d_lo,
                                            int r, int t, const Categorical &cat, int shift) {
    using G = Geo<S>;
    constexpr int NB = (3 * S + 7) / 8;
#pragma unroll
    for (int w = 0; w < G::TP / 4; w++) words[w] = 0;
    uint32_t nz[3] = {0, 0, 0};
#pragma unroll
    for (int bq = 0; bq < NB; bq++) {
        uint32_t blk[4];
        philox4x32_10((uint32_t)bq, (uint32_t)t, (uint32_t)r, d_lo, k0, k1, blk);
#pragma unroll
        for (int h = 0; h < 8; h++) {
            const int q = bq * 8 + h;
            if (q < 3 * S) {
                const uint32_t x = (h & 1) ? (blk[h >> 1] >> 16) : (blk[h >> 1] & 0xFFFFu);
                int idx = 0;
#pragma unroll
                for (int i = 0; i < 7; i++) idx += (i < cat.n - 1 && x >= cat.thr[i]) ? 1 : 0;
                int val = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) val = (idx == i) ? (int)cat.values[i] : val;
                nz[q / S] |= (uint32_t)(val != 0);
                words[q >> 2] |= (uint32_t)((val + shift) & 0xFF) << (8 * (q & 3));
            }
        }
    }
    return (nz[0] & nz[1] & nz[2]) != 0;
}

template <int S, int NT, int NPASS, bool SAMPLE>
__global__ void __launch_bounds__(NT)
    demo_kernel(unsigned long long seed, unsigned long long first_demo, long long N, int R, int shift, Categorical cat,
                int max_tries, int chk, uint8_t *__restrict__ tape, long long tape_step_stride, int8_t *__restrict__ slab,
                uint8_t *__restrict__ flags) {
    using C = DemoCfg<S, NT, NPASS>;
    using G = Geo<S>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_slab = smem;
    uint8_t *s_tape = smem + C::SLAB_BYTES;                                  // [R][TG][TP]
    uint32_t *s_part = reinterpret_cast<uint32_t *>(s_tape + C::tape_bytes(R)); // [TG][PW]
    uint32_t *s_flag = s_part + C::TG * C::PW;                                // [TG]
    uint32_t *s_work = s_flag + C::TG;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_work + 2);

    const int tid = threadIdx.x;
    const long long g0 = (long long)blockIdx.x * C::TG;
    const int ng = (int)min((long long)C::TG, N - g0);

    for (int w = tid; w < C::SLAB_BYTES / 4; w += NT) reinterpret_cast<uint32_t *>(s_slab)[w] = 0;
    for (int g = tid; g < C::TG; g += NT) s_flag[g] = 0;
    if (tid == 0) {
        *s_work = 0;
        if constexpr (!SAMPLE) {
            mbar_init(s_bar, 1);
            mbar_fence_init();
        }
    }
    __syncthreads();

    if constexpr (SAMPLE) {
        // ---------------- A. draw the factor triples of the tile
        const int npairs = ng * R;
        for (;;) {
            const int p = (int)atomicAdd(s_work, 1u);
            if (p >= npairs) break;
            const int g = p / R, r = p - g * R;
            const unsigned long long d = first_demo + (unsigned long long)(g0 + g);
            const uint32_t k0 = (uint32_t)seed ^ ((uint32_t)(d >> 32) * 0x9E3779B9u), k1 = (uint32_t)(seed >> 32);
            uint32_t words[G::TP / 4];
            bool ok = false;
            for (int t = 0; t < max_tries && !ok; t++) ok = draw_triple<S>(words, k0, k1, (uint32_t)d, r, t, cat, shift);
            if (!ok) { // bounded retries: forced unit triple (the reference would loop forever)
#pragma unroll
                for (int w = 0; w < G::TP / 4; w++) words[w] = 0;
                const int top = (int)cat.values[cat.n - 1] + shift;
#pragma unroll
                for (int q = 0; q < 3 * S; q++)
                    words[q >> 2] |= (uint32_t)(((q % S) == 0 ? top : shift) & 0xFF) << (8 * (q & 3));
                atomicOr(&s_flag[g], 8u);
            }
            uint32_t *dst = reinterpret_cast<uint32_t *>(s_tape + ((size_t)r * C::TG + g) * G::TP);
#pragma unroll
            for (int w = 0; w < G::TP / 4; w++) dst[w] = words[w];
        }
    } else {
        // ---------------- A'. replay: bulk-load the tape tile [R][ng][TP]
        if (tid == 0) {
            mbar_expect_tx(s_bar, (uint32_t)(R * ng * G::TP));
            for (int r = 0; r < R; r++)
                bulk_g2s(s_tape + (size_t)r * C::TG * G::TP, tape + (size_t)r * tape_step_stride + g0 * G::TP,
                         (uint32_t)(ng * G::TP), s_bar);
        }
        mbar_wait(s_bar, 0);
    }
    __syncthreads();

    // ---------------- B. accumulate the R rank-1 terms in registers
    Lane<S> L;
    const bool active = tid < C::ACTIVE;
    const int gl = tid / G::WR;
    L.init(active ? tid % G::WR : 0);
#pragma unroll 1
    for (int p = 0; p < NPASS; p++) {
        const int g = p * C::GPASS + gl;
        uint32_t pr = 0;
        if (active && g < ng) {
            int32_t acc[S];
#pragma unroll
            for (int i = 0; i < S; i++) acc[i] = 0;
            uint32_t bad = 0;
            int until = chk;
            for (int r = 0; r < R; r++) {
                const uint8_t *tok = s_tape + ((size_t)r * C::TG + g) * G::TP;
                const int32_t vw = pack_vw<S>(tok, L, shift);
                const uint4 ut = *reinterpret_cast<const uint4 *>(tok);
                const uint32_t uw[4] = {ut.x, ut.y, ut.z, ut.w};
#pragma unroll
                for (int i = 0; i < S; i++) acc[i] += ((int)((uw[i >> 2] >> (8 * (i & 3))) & 0xFFu) - shift) * vw;
                if (--until == 0 || r == R - 1) { // every entry still in [-64,63]? then the next chk terms cannot alias
                    until = chk;
#pragma unroll
                    for (int i = 0; i < S; i++) {
                        const uint32_t ob = (uint32_t)acc[i] + H4;
                        bad |= ~(ob ^ (ob << 1));
                    }
                }
            }
            uint32_t cnt = 0;
            uint32_t *col = reinterpret_cast<uint32_t *>(s_slab + (size_t)g * G::GP) + L.c;
#pragma unroll
            for (int i = 0; i < S; i++) {
                const uint32_t t = ((uint32_t)acc[i] + H4) ^ H4;
                col[i * G::WR] = t;
                cnt += ((((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & L.hv) >> 7;
            }
            pr = make_partial(byte_sum(cnt), true, (bad & L.hv) != 0);
        }
        if ((pr >> 24) != 0) atomicOr(&s_flag[g], (uint32_t)TG_FLAG_RANGE);
    }
    fence_proxy_async();
    __syncthreads();

    // ---------------- C. tile out
    if (tid == 0) {
        bulk_s2g(slab + g0 * G::GP, s_slab, (uint32_t)(ng * G::GP));
        if constexpr (SAMPLE) {
            for (int r = 0; r < R; r++)
                bulk_s2g(tape + (size_t)r * tape_step_stride + g0 * G::TP, s_tape + (size_t)r * C::TG * G::TP,
                         (uint32_t)(ng * G::TP));
        }
        bulk_commit();
    }
    if (flags)
        for (int g = tid; g < ng; g += NT) flags[g0 + g] = (uint8_t)s_flag[g];
    if (tid == 0) bulk_wait<0>();
}

template <int S, int NT, int NPASS, bool SAMPLE>
static int launch_demo(unsigned long long seed, unsigned long long first, long long N, int R, int shift,
                       const Categorical &cat, int max_tries, uint8_t *tape, long long stride, int8_t *slab,
                       uint8_t *flags, cudaStream_t st) {
    using C = DemoCfg<S, NT, NPASS>;
    auto kern = demo_kernel<S, NT, NPASS, SAMPLE>;
    const int smem = C::smem_bytes(R);
    if (smem > 227 * 1024) return TG_E_ARG;
    TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int s3 = shift * shift * shift;
    const int chk = s3 >= 64 ? 1 : 64 / s3;
    const long long grid = (N + C::TG - 1) / C::TG;
    if (grid > 0x7FFFFFFFLL) return TG_E_ARG;
    kern<<<(int)grid, NT, smem, st>>>(seed, first, N, R, shift, cat, max_tries, chk, tape, stride, slab, flags);
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

template <bool SAMPLE>
static int dispatch_demo(unsigned long long seed, unsigned long long first, long long N, int R, int S, int shift,
                         const Categorical &cat, int max_tries, uint8_t *tape, long long stride, int8_t *slab,
                         uint8_t *flags, cudaStream_t st) {
    switch (S) {
    case 4: return launch_demo<4, 256, 2, SAMPLE>(seed, first, N, R, shift, cat, max_tries, tape, stride, slab, flags, st);
    case 9: return launch_demo<9, 256, 2, SAMPLE>(seed, first, N, R, shift, cat, max_tries, tape, stride, slab, flags, st);
    case 16: return launch_demo<16, 256, 1, SAMPLE>(seed, first, N, R, shift, cat, max_tries, tape, stride, slab, flags, st);
    }
    return TG_E_ARG;
}

} // namespace tg

extern "C" {

int tg_demo_gen_philox(uint64_t seed, uint64_t first_demo, int64_t N, int R, int S, int shift, const int8_t *values,
                       const double *probs, int n_values, int max_tries, uint8_t *tape, int64_t tape_step_stride,
                       int8_t *slab, uint8_t *flags, void *stream) {
    if (!tg::supported_S(S) || N < 0 || R < 1 || shift < 1 || shift > 4 || n_values < 1 || n_values > 8 || max_tries < 1)
        return TG_E_ARG;
    if (N == 0) return TG_OK;
    if (!values || !probs || !tape || !slab) return TG_E_ARG;
    if (((uintptr_t)tape | (uintptr_t)slab | (uintptr_t)tape_step_stride) & 15) return TG_E_ARG;
    tg::Categorical cat;
    double total = 0, run = 0;
    for (int i = 0; i < n_values; i++) {
        if (!(probs[i] >= 0) || values[i] < -shift || values[i] > shift) return TG_E_ARG;
        total += probs[i];
    }
    if (!(total > 0)) return TG_E_ARG;
    for (int i = 0; i < 8; i++) {
        cat.thr[i] = 65536u;
        cat.values[i] = 0;
    }
    for (int i = 0; i < n_values; i++) {
        run += probs[i] / total;
        const double t = run * 65536.0;
        cat.thr[i] = (i == n_values - 1 || t >= 65536.0) ? 65536u : (uint32_t)t;
        cat.values[i] = values[i];
    }
    cat.n = n_values;
    return tg::dispatch_demo<true>(seed, first_demo, N, R, S, shift, cat, max_tries, tape, tape_step_stride, slab, flags,
                                   (cudaStream_t)stream);
}

int tg_demo_accumulate(const uint8_t *tape, int64_t tape_step_stride, int64_t N, int R, int S, int shift, int8_t *slab,
                       uint8_t *flags, void *stream) {
    if (!tg::supported_S(S) || N < 0 || R < 1 || shift < 1 || shift > 4) return TG_E_ARG;
    if (N == 0) return TG_OK;
    if (!tape || !slab) return TG_E_ARG;
    if (((uintptr_t)tape | (uintptr_t)slab | (uintptr_t)tape_step_stride) & 15) return TG_E_ARG;
    tg::Categorical cat = {};
    return tg::dispatch_demo<false>(0, 0, N, R, S, shift, cat, 1, const_cast<uint8_t *>(tape), tape_step_stride, slab,
                                    flags, (cudaStream_t)stream);
}

} // extern "C"
