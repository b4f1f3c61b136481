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
//   B. S threads per demo, thread j owning the entries (i, j, 0..S-1) of every
//      row i in registers as packed words: per term  c = u_i * v_j  and
//      acc[i][m] += c * pack(w[4m..4m+3])  -- one IMAD per four entries, the
//      pack(w) words shared by all (i, j) (packed arithmetic of tg_step.cuh),
//      with a range check often enough that the packed form can never alias.
//   C. the tile (slab + tape) leaves through TMA bulk stores.
#include "tg_step.cuh"

namespace tg {

struct Categorical {
    uint32_t thr[8];   // 16-bit CDF thresholds (65536 = never exceeded)
    uint32_t lut_lo;   // token (value + shift) of buckets 0-3, one byte each
    uint32_t lut_hi;   // buckets 4-7
    uint32_t zero_pat; // token of the value 0 in every byte (0xFFFFFFFF if 0 is not in the alphabet)
    uint32_t top_tok;  // token of the last bucket (forced unit triple)
    int n;
};

// one Philox4x32-10 block; IMAD.WIDE gives hi and lo of each product in one instruction
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        c1 = (uint32_t)p1;
        c2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

template <int S, int NT, int NPASS>
struct DemoCfg {
    using G = Geo<S>;
    static constexpr int GPASS = NT / S;  // S threads per demo (thread = factor index j)
    static constexpr int TG = GPASS * NPASS;
    static constexpr int ACTIVE = GPASS * S;
    static constexpr int KW = (S + 3) / 4; // packed words per (i, j) run of S entries
    static constexpr int SLAB_BYTES = TG * G::GP;
    static __host__ __device__ constexpr int tape_bytes(int R) { return R * TG * G::TP; }
    static __host__ __device__ constexpr int smem_bytes(int R) { return SLAB_BYTES + tape_bytes(R) + TG * 4 + 16; }
};

// "is this factor all zero" over packed token words: OR of (word ^ zero_pat) under the factor's byte mask
template <int S>
__device__ __forceinline__ bool factor_nonzero(const uint32_t words[Geo<S>::TP / 4], int f, uint32_t zero_pat) {
    uint32_t acc = 0;
#pragma unroll
    for (int w = 0; w < Geo<S>::TP / 4; w++) {
        uint32_t m = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int q = 4 * w + b;
            if (q >= f * S && q < (f + 1) * S) m |= 0xFFu << (8 * b);
        }
        if (m) acc |= (words[w] ^ zero_pat) & m;
    }
    return acc != 0;
}

// draw one factor triple (3S tokens) for (demo key k0/k1/d_lo, term r, try t); true if accepted.
// Draw q is the 16-bit half (q & 1) of word (q >> 1) & 3 of Philox block q >> 3; bucket = number of
// thresholds <= draw; four bucket indexes form a PRMT selector that looks the four tokens up at once.
template <int S, int NTHR>
__device__ __forceinline__ bool draw_triple(uint32_t words[Geo<S>::TP / 4], uint32_t k0, uint32_t k1, uint32_t d_lo,
                                            int r, int t, const Categorical &cat) {
    using G = Geo<S>;
    constexpr int NB = (3 * S + 7) / 8;
#pragma unroll
    for (int w = 0; w < G::TP / 4; w++) words[w] = 0;
#pragma unroll
    for (int bq = 0; bq < NB; bq++) {
        uint32_t blk[4];
        philox4x32_10((uint32_t)bq, (uint32_t)t, (uint32_t)r, d_lo, k0, k1, blk);
#pragma unroll
        for (int half = 0; half < 2; half++) { // tokens 8bq + 4half .. +3  ->  token word 2bq + half
            if (8 * bq + 4 * half < 3 * S) {
                uint32_t sel = 0;
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const uint32_t wv = blk[2 * half + (h >> 1)];
                    const uint32_t x = (h & 1) ? (wv >> 16) : (wv & 0xFFFFu);
#pragma unroll
                    for (int i = 0; i < NTHR; i++) sel += (x >= cat.thr[i]) ? (1u << (4 * h)) : 0u;
                }
                uint32_t tokw = __byte_perm(cat.lut_lo, cat.lut_hi, sel);
                const int q0 = 8 * bq + 4 * half;
                if (q0 + 4 > 3 * S) tokw &= 0xFFFFFFFFu >> (8 * (q0 + 4 - 3 * S)); // tape padding stays zero
                words[2 * bq + half] = tokw;
            }
        }
    }
    return factor_nonzero<S>(words, 0, cat.zero_pat) && factor_nonzero<S>(words, 1, cat.zero_pat) &&
           factor_nonzero<S>(words, 2, cat.zero_pat);
}

template <int S, int NT, int NPASS, bool SAMPLE, int NTHR>
__global__ void __launch_bounds__(NT, S == 16 ? 2 : 3)
    demo_kernel(unsigned long long seed, unsigned long long first_demo, long long N, int R, int shift, Categorical cat,
                int max_tries, int chk, uint8_t *__restrict__ tape, long long tape_step_stride, int8_t *__restrict__ slab,
                uint8_t *__restrict__ flags) {
    using C = DemoCfg<S, NT, NPASS>;
    using G = Geo<S>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_slab = smem;
    uint8_t *s_tape = smem + C::SLAB_BYTES;                                  // [R][TG][TP]
    uint32_t *s_flag = reinterpret_cast<uint32_t *>(s_tape + C::tape_bytes(R)); // [TG]
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
        // ---------------- A. draw the factor triples of the tile.  Every lane runs a small state machine
        // (pair, try): a rejected triple just bumps the try, an accepted one is stored and the lane claims the
        // next pair from the shared counter -- so lanes never wait for another lane's rejection loop.
        const int npairs = ng * R;
        int p = (int)atomicAdd(s_work, 1u), t = 0;
        int g = 0, r = 0;
        uint32_t k0 = 0, d_lo = 0;
        const uint32_t k1 = (uint32_t)(seed >> 32);
        auto claim = [&]() {
            if (p < npairs) {
                g = p / R, r = p - g * R;
                const unsigned long long d = first_demo + (unsigned long long)(g0 + g);
                k0 = (uint32_t)seed ^ ((uint32_t)(d >> 32) * 0x9E3779B9u), d_lo = (uint32_t)d;
            }
        };
        claim();
        while (p < npairs) {
            uint32_t words[G::TP / 4];
            bool ok = draw_triple<S, NTHR>(words, k0, k1, d_lo, r, t, cat);
            if (!ok && t + 1 >= max_tries) { // bounded retries: forced unit triple (the reference would loop forever, Q11)
#pragma unroll
                for (int w = 0; w < G::TP / 4; w++) words[w] = 0;
#pragma unroll
                for (int q = 0; q < 3 * S; q++)
                    words[q >> 2] |= (((q % S) == 0 ? cat.top_tok : (uint32_t)shift) & 0xFFu) << (8 * (q & 3));
                atomicOr(&s_flag[g], 8u);
                ok = true;
            }
            if (ok) {
                uint32_t *dst = reinterpret_cast<uint32_t *>(s_tape + ((size_t)r * C::TG + g) * G::TP);
#pragma unroll
                for (int w = 0; w < G::TP / 4; w++) dst[w] = words[w];
                p = (int)atomicAdd(s_work, 1u), t = 0;
                claim();
            } else {
                t++;
            }
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
    constexpr int KW = C::KW;
    const bool active = tid < C::ACTIVE;
    const int gl = tid / S, j = tid % S;
    const int vword = ((S + j) >> 2) * 4;                 // aligned word of the tape holding v_j
    const uint32_t vhot = 1u << (8 * ((S + j) & 3));      // one-hot selector of v_j inside it
    constexpr uint32_t WLAST = (S % 4) ? (0xFFFFFFFFu >> (8 * (4 - S % 4))) : 0xFFFFFFFFu; // valid bytes of the last w word
#pragma unroll 1
    for (int p = 0; p < NPASS; p++) {
        const int g = p * C::GPASS + gl;
        if (active && g < ng) {
            int32_t acc[S][KW];
#pragma unroll
            for (int i = 0; i < S; i++)
#pragma unroll
                for (int m = 0; m < KW; m++) acc[i][m] = 0;
            uint32_t bad = 0;
            int until = chk;
            for (int r = 0; r < R; r++) {
                const uint8_t *tok = s_tape + ((size_t)r * C::TG + g) * G::TP;
                const uint32_t *tw = reinterpret_cast<const uint32_t *>(tok);
                // packed w coefficients, bytes 2S .. 3S-1 of the tape record (funnel shift when not word aligned)
                int32_t wp[KW];
#pragma unroll
                for (int m = 0; m < KW; m++) {
                    constexpr int o = 2 * S;
                    const uint32_t lo = tw[(o >> 2) + m];
                    uint32_t wt = lo;
                    if constexpr ((o & 3) != 0) wt = __funnelshift_r(lo, tw[(o >> 2) + m + 1], 8 * (o & 3));
                    const uint32_t msk = (m == KW - 1) ? WLAST : 0xFFFFFFFFu;
                    wp[m] = (int32_t)((wt & msk) - (((uint32_t)shift * ONES4) & msk));
                }
                const int vj = (int)__dp4a(*reinterpret_cast<const uint32_t *>(tok + vword), vhot, (uint32_t)(-shift));
                const uint4 ut = *reinterpret_cast<const uint4 *>(tok);
                const uint32_t uw[4] = {ut.x, ut.y, ut.z, ut.w};
#pragma unroll
                for (int i = 0; i < S; i++) {
                    const int cij = coef_u(uw, i, shift) * vj;
#pragma unroll
                    for (int m = 0; m < KW; m++) acc[i][m] += cij * wp[m];
                }
                if (--until == 0 || r == R - 1) { // every entry still in [-64,63]? then the next chk terms cannot alias
                    until = chk;
#pragma unroll
                    for (int i = 0; i < S; i++)
#pragma unroll
                        for (int m = 0; m < KW; m++) {
                            const uint32_t ob = (uint32_t)acc[i][m] + H4;
                            bad |= ~(ob ^ (ob << 1)) & ((m == KW - 1) ? (WLAST & H4) : H4);
                        }
                }
            }
            // registers -> slab tile: entry (i, j, k) is byte i*RP + j*S + k
            uint8_t *gbase = s_slab + (size_t)g * G::GP + j * S;
#pragma unroll
            for (int i = 0; i < S; i++) {
                if constexpr (S % 4 == 0) {
#pragma unroll
                    for (int m = 0; m < KW; m++)
                        reinterpret_cast<uint32_t *>(gbase + i * G::RP)[m] = ((uint32_t)acc[i][m] + H4) ^ H4;
                } else {
#pragma unroll
                    for (int k = 0; k < S; k++) {
                        const uint32_t t = ((uint32_t)acc[i][k >> 2] + H4) ^ H4;
                        gbase[i * G::RP + k] = (uint8_t)(t >> (8 * (k & 3)));
                    }
                }
            }
            if (bad) atomicOr(&s_flag[g], (uint32_t)TG_FLAG_RANGE);
        }
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

template <int S, int NT, int NPASS, bool SAMPLE, int NTHR>
static int launch_demo(unsigned long long seed, unsigned long long first, long long N, int R, int shift,
                       const Categorical &cat, int max_tries, uint8_t *tape, long long stride, int8_t *slab,
                       uint8_t *flags, cudaStream_t st) {
    using C = DemoCfg<S, NT, NPASS>;
    auto kern = demo_kernel<S, NT, NPASS, SAMPLE, NTHR>;
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

template <bool SAMPLE, int NTHR>
static int dispatch_demo(unsigned long long seed, unsigned long long first, long long N, int R, int S, int shift,
                         const Categorical &cat, int max_tries, uint8_t *tape, long long stride, int8_t *slab,
                         uint8_t *flags, cudaStream_t st) {
    switch (S) {
    case 4: return launch_demo<4, 256, 1, SAMPLE, NTHR>(seed, first, N, R, shift, cat, max_tries, tape, stride, slab, flags, st);
    case 9: return launch_demo<9, 256, 1, SAMPLE, NTHR>(seed, first, N, R, shift, cat, max_tries, tape, stride, slab, flags, st);
    case 16: return launch_demo<16, 256, 1, SAMPLE, NTHR>(seed, first, N, R, shift, cat, max_tries, tape, stride, slab, flags, st);
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
    tg::Categorical cat = {};
    double total = 0, run = 0;
    for (int i = 0; i < n_values; i++) {
        if (!(probs[i] >= 0) || values[i] < -shift || values[i] > shift) return TG_E_ARG;
        total += probs[i];
    }
    if (!(total > 0)) return TG_E_ARG;
    uint8_t lut[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cat.zero_pat = 0xFFFFFFFFu;
    for (int i = 0; i < 8; i++) cat.thr[i] = 65536u;
    for (int i = 0; i < n_values; i++) {
        run += probs[i] / total;
        const double t = run * 65536.0;
        cat.thr[i] = (i == n_values - 1 || t >= 65536.0) ? 65536u : (uint32_t)t;
        lut[i] = (uint8_t)(values[i] + shift);
        if (values[i] == 0) cat.zero_pat = (uint32_t)shift * 0x01010101u;
    }
    cat.lut_lo = lut[0] | (lut[1] << 8) | (lut[2] << 16) | ((uint32_t)lut[3] << 24);
    cat.lut_hi = lut[4] | (lut[5] << 8) | (lut[6] << 16) | ((uint32_t)lut[7] << 24);
    cat.top_tok = lut[n_values - 1];
    cat.n = n_values;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_values <= 3)
        return tg::dispatch_demo<true, 2>(seed, first_demo, N, R, S, shift, cat, max_tries, tape, tape_step_stride, slab, flags, st);
    if (n_values <= 5)
        return tg::dispatch_demo<true, 4>(seed, first_demo, N, R, S, shift, cat, max_tries, tape, tape_step_stride, slab, flags, st);
    return tg::dispatch_demo<true, 7>(seed, first_demo, N, R, S, shift, cat, max_tries, tape, tape_step_stride, slab, flags, st);
}

int tg_demo_accumulate(const uint8_t *tape, int64_t tape_step_stride, int64_t N, int R, int S, int shift, int8_t *slab,
                       uint8_t *flags, void *stream) {
    if (!tg::supported_S(S) || N < 0 || R < 1 || shift < 1 || shift > 4) return TG_E_ARG;
    if (N == 0) return TG_OK;
    if (!tape || !slab) return TG_E_ARG;
    if (((uintptr_t)tape | (uintptr_t)slab | (uintptr_t)tape_step_stride) & 15) return TG_E_ARG;
    tg::Categorical cat = {};
    return tg::dispatch_demo<false, 2>(0, 0, N, R, S, shift, cat, 1, const_cast<uint8_t *>(tape), tape_step_stride, slab,
                                       flags, (cudaStream_t)stream);
}

} // extern "C"
