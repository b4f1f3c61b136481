// tg_aux.cu -- K4 demo_sample (training-sample batcher), K6 slice_rank, K7 state_key.
#include <type_traits>

#include "tg_step.cuh"

namespace tg {

// ------------------------------------------------------------------ K4
// SyntheticDemoDataset.__getitem__ (datasets.py:77-122) for a batch of sample
// indices, straight from the in-HBM demo store (no torch.load, no Python
// replay).  idx = demo * R + action.  Emits, per sample,
//   state  float32 [dim_t][S][S][S]: slot 0 = target - sum_{j>a} rank1(tok_j),
//          slots 1.. = rank1 of actions min(a+dim_t-1,R-1) .. a+1, zero padded
//   scalar float32 = R - a, reward float32 = -(a+1), action int64 [3S] = tok_a.
// rank1 uses coefficient = token - replay_shift; the reference hard-codes 1
// there, callers wanting the true residual pass the demo's shift.
// One thread per (sample, word column); per-entry int32 accumulators, so any
// magnitude the float32 reference can hold exactly is exact here too.
template <int S, typename TB>
__device__ __forceinline__ void vw_bytes(const TB *tok, const Lane<S> &L, int shift, int vw[4]) {
    const int vA = (int)tok[L.off_vA] - shift, vB = (int)tok[L.off_vB] - shift;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const bool inA = (L.maskA >> (8 * b)) & 1u, inB = (L.maskB >> (8 * b)) & 1u;
        const int w = (int)tok[L.off_w[b]] - shift;
        vw[b] = inA ? vA * w : (inB ? vB * w : 0);
    }
}

// sign-extended byte B of a word (one PRMT with a sign-replicating selector)
template <int B>
__device__ __forceinline__ int sext_byte(uint32_t w) {
    constexpr uint32_t sel = (uint32_t)B | ((uint32_t)(B | 8) << 4) | ((uint32_t)(B | 8) << 8) | ((uint32_t)(B | 8) << 12);
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0u), "r"(sel));
    return (int)d;
}

// sample index -> (demo, action); indices below 2^32 (every realistic store) take a 32-bit division
__device__ __forceinline__ void split_index(long long id, int R, long long &demo, int &a) {
    if ((unsigned long long)id < 0x100000000ULL) {
        const uint32_t d = (uint32_t)id / (uint32_t)R;
        demo = (long long)d;
        a = (int)((uint32_t)id - d * (uint32_t)R);
    } else {
        demo = id / R;
        a = (int)(id - demo * R);
    }
}

// PACK16: the head is accumulated two entries per 32-bit word (16-bit lanes in integer form, one IMAD per two
// entries, half the registers -> more samples in flight); the host takes this path when R * cmax^3 + 128 fits int16.
// STAGE: the CTA first copies the action records a .. R-1 of all its samples into shared memory, every load in flight at
// once (the step-major tape puts each record of a demo in a different DRAM page; read one after the other inside the
// replay loop they serialise into R - a round trips per sample, which is what bounded this kernel).
template <int S, bool PACK16, bool STAGE>
__global__ void __launch_bounds__(128, PACK16 ? (S == 16 ? 6 : 8) : 1)
    demo_sample_kernel(const uint8_t *__restrict__ tape, long long tape_step_stride,
                                   const int8_t *__restrict__ slab, long long N, int R, int dim_t, int replay_shift,
                                   const long long *__restrict__ idx, long long nb, float *__restrict__ states,
                                   float *__restrict__ scalars, long long *__restrict__ actions,
                                   float *__restrict__ rewards) {
    using G = Geo<S>;
    // COEF (the production variant): the staged records after the sample's own action hold int8 coefficients; exact under the
    // same contract PACK16 already needs (tokens <= 8, 0 <= replay_shift <= 8)
    constexpr bool COEF = PACK16 && STAGE;
    extern __shared__ __align__(16) uint8_t s_tok[]; // STAGE: [samples of the CTA][R][TP]
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long b = t / G::WR;
    const long long b0 = (blockIdx.x * (long long)blockDim.x) / G::WR; // first sample this CTA touches
    // this thread's own sample first: its index and the words of the target it owns are in flight while the CTA stages the
    // action records (otherwise a third dependent DRAM round trip after the barrier)
    Lane<S> L;
    L.init((int)(t % G::WR));
    long long demo = -1;
    int a = 0;
    if (b < nb) split_index(idx[b], R, demo, a);
    const bool live = b < nb && demo >= 0 && demo < N;
    uint32_t head[S];
    if (live) {
        const int8_t *tg = slab + demo * G::GP + 4 * L.c;
#pragma unroll
        for (int i = 0; i < S; i++) head[i] = __ldg(reinterpret_cast<const uint32_t *>(tg + i * G::RP));
    }
    if constexpr (STAGE) {
        constexpr int SLOTS = (128 + G::WR - 1) / G::WR + 1;
        const int nslots = (int)min((long long)SLOTS, nb - b0);
        for (int item = threadIdx.x; item < nslots * R; item += 128) {
            const int sl = item / R, j = item - sl * R;
            long long demo2;
            int a2;
            split_index(idx[b0 + sl], R, demo2, a2);
            if (demo2 >= 0 && demo2 < N && j >= a2) {
                const uint4 *src = reinterpret_cast<const uint4 *>(tape + (size_t)j * tape_step_stride + demo2 * G::TP);
                uint4 *dst = reinterpret_cast<uint4 *>(s_tok + ((size_t)sl * R + j) * G::TP);
                if (COEF && j > a2) {
                    // records that are only ever replayed are staged as int8 coefficients (token - replay_shift, no borrow
                    // between bytes: token | 0x80 > shift), so the replay loop needs no subtraction per byte
                    const uint32_t sh4 = (uint32_t)replay_shift * ONES4;
#pragma unroll
                    for (int w = 0; w < G::TP / 16; w++) {
                        const uint4 x = __ldg(src + w);
                        dst[w] = make_uint4(((x.x | H4) - sh4) ^ H4, ((x.y | H4) - sh4) ^ H4, ((x.z | H4) - sh4) ^ H4, ((x.w | H4) - sh4) ^ H4);
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < G::TP / 16; w++) dst[w] = __ldg(src + w);
                }
            }
        }
        __syncthreads();
    }
    if (!live) return;
    const uint8_t *tk = STAGE ? s_tok + (size_t)(b - b0) * R * G::TP : tape + demo * G::TP;
    if constexpr (STAGE) tape_step_stride = G::TP;
    // head
    float *st = states + b * (long long)dim_t * G::S3;
    const int nv = min(4, G::S2 - 4 * L.c); // entries of this word column inside a row (the last column of 9x9x9 holds one)
    if constexpr (PACK16) {
        int acc[S][2];
#pragma unroll
        for (int i = 0; i < S; i++) {
            const uint32_t t = head[i];
            acc[i][0] = (int)(int8_t)(t & 0xFFu) + ((int)(int8_t)((t >> 8) & 0xFFu)) * 65536;
            acc[i][1] = (int)(int8_t)((t >> 16) & 0xFFu) + ((int)(int8_t)(t >> 24)) * 65536;
        }
        if constexpr (COEF) {
            // entries outside the row (last word column of 9x9x9) pick up a v of their own; they are never stored and a low
            // 16-bit lane does not depend on the lane above it
            const bool inA0 = !G::STRADDLE || (L.maskA & 0x1u), inA1 = !G::STRADDLE || (L.maskA & 0x100u),
                       inA2 = !G::STRADDLE || (L.maskA & 0x10000u), inA3 = !G::STRADDLE || (L.maskA & 0x1000000u);
            const int8_t *rec = reinterpret_cast<const int8_t *>(tk) + (size_t)(a + 1) * G::TP;
#pragma unroll 2
            for (int j = a + 1; j < R; j++, rec += G::TP) {
                const int nvA = -(int)rec[L.off_vA], nvB = G::STRADDLE ? -(int)rec[L.off_vB] : nvA;
                const int w0 = rec[L.off_w[0]], w1 = rec[L.off_w[1]], w2 = rec[L.off_w[2]], w3 = rec[L.off_w[3]];
                const int np0 = (inA0 ? nvA : nvB) * w0 + ((inA1 ? nvA : nvB) * w1) * 65536;
                const int np1 = (inA2 ? nvA : nvB) * w2 + ((inA3 ? nvA : nvB) * w3) * 65536;
                uint32_t uq[4];
                if constexpr (S <= 4) {
                    uq[0] = *reinterpret_cast<const uint32_t *>(rec);
                } else {
                    const uint4 q4 = *reinterpret_cast<const uint4 *>(rec);
                    uq[0] = q4.x, uq[1] = q4.y, uq[2] = q4.z, uq[3] = q4.w;
                }
#pragma unroll
                for (int i = 0; i < S; i++) {
                    int u;
                    switch (i & 3) {
                    case 0: u = sext_byte<0>(uq[i >> 2]); break;
                    case 1: u = sext_byte<1>(uq[i >> 2]); break;
                    case 2: u = sext_byte<2>(uq[i >> 2]); break;
                    default: u = sext_byte<3>(uq[i >> 2]); break;
                    }
                    acc[i][0] += u * np0;
                    acc[i][1] += u * np1;
                }
            }
        } else {
            for (int j = a + 1; j < R; j++) {
                const uint8_t *tok = tk + (size_t)j * tape_step_stride;
                int vw[4];
                vw_bytes<S, uint8_t>(tok, L, replay_shift, vw);
                const int p0 = vw[0] + vw[1] * 65536, p1 = vw[2] + vw[3] * 65536;
#pragma unroll
                for (int i = 0; i < S; i++) {
                    const int u = (int)tok[i] - replay_shift;
                    acc[i][0] -= u * p0;
                    acc[i][1] -= u * p1;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < S; i++) {
            float f[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int x = acc[i][q >> 1];
                const int lo = (int)(short)(x & 0xFFFF);
                f[q] = (float)((q & 1) ? (x - lo) >> 16 : lo);
            }
            store_run<S>(st + i * G::S2 + 4 * L.c, nv, f[0], f[1], f[2], f[3]);
        }
    } else {
        int acc[S][4];
#pragma unroll
        for (int i = 0; i < S; i++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[i][q] = (int)(int8_t)((head[i] >> (8 * q)) & 0xFFu);
        for (int j = a + 1; j < R; j++) {
            const uint8_t *tok = tk + (size_t)j * tape_step_stride;
            int vw[4];
            vw_bytes<S, uint8_t>(tok, L, replay_shift, vw);
#pragma unroll
            for (int i = 0; i < S; i++) {
                const int u = (int)tok[i] - replay_shift;
#pragma unroll
                for (int q = 0; q < 4; q++) acc[i][q] -= u * vw[q];
            }
        }
#pragma unroll
        for (int i = 0; i < S; i++)
            store_run<S>(st + i * G::S2 + 4 * L.c, nv, (float)acc[i][0], (float)acc[i][1], (float)acc[i][2], (float)acc[i][3]);
    }
    // history slots: rank-1 tensors of the next actions, latest first
    const int hi = min(a + dim_t, R);
    for (int s = 1; s < dim_t; s++) {
        const int j = hi - s;
        float *ss = st + (long long)s * G::S3;
        if (j >= a + 1) {
            using TB = typename std::conditional<COEF, int8_t, uint8_t>::type; // COEF: the record already holds coefficients
            const int hshift = COEF ? 0 : replay_shift;
            const TB *tok = reinterpret_cast<const TB *>(tk + (size_t)j * tape_step_stride);
            int vw[4];
            vw_bytes<S, TB>(tok, L, hshift, vw);
#pragma unroll
            for (int i = 0; i < S; i++) {
                const int u = (int)tok[i] - hshift;
                store_run<S>(ss + i * G::S2 + 4 * L.c, nv, (float)(u * vw[0]), (float)(u * vw[1]), (float)(u * vw[2]), (float)(u * vw[3]));
            }
        } else {
#pragma unroll
            for (int i = 0; i < S; i++) store_run<S>(ss + i * G::S2 + 4 * L.c, nv, 0.f, 0.f, 0.f, 0.f);
        }
    }
    if (L.c == 0) {
        scalars[b] = (float)(R - a);
        rewards[b] = -(float)(a + 1);
    }
    const uint8_t *ta = tk + (size_t)a * tape_step_stride;
    for (int q = L.c; q < 3 * S; q += G::WR) actions[b * 3 * S + q] = (long long)ta[q];
}

// ------------------------------------------------------------------ K6
// get_rank (utils.py:134-140): sum over the S slices T[i,:,:] of the matrix
// rank.  The reference takes a float32 SVD; here the rank is computed exactly
// over GF(p), p = 2^31-1, by fraction-free elimination (row <- row*piv -
// row[c]*pivrow), one matrix per group of S lanes, lane = matrix row.  rank_p
// <= rank_Q with equality unless p divides a pivot minor (probability ~S/p per
// matrix); the small integer matrices of the game are far from that.
__device__ __forceinline__ uint32_t mulmod31(uint32_t a, uint32_t b) {
    const unsigned long long z = (unsigned long long)a * b;
    uint32_t r = (uint32_t)(z & 0x7FFFFFFFu) + (uint32_t)(z >> 31);
    r = (r & 0x7FFFFFFFu) + (r >> 31);
    return r >= 0x7FFFFFFFu ? r - 0x7FFFFFFFu : r;
}

template <int S>
__global__ void slice_rank_kernel(const int8_t *__restrict__ slab, int32_t *__restrict__ ranks, long long B) {
    using G = Geo<S>;
    constexpr int LG = S <= 4 ? 4 : (S <= 8 ? 8 : (S <= 16 ? 16 : 32)); // lanes per matrix (power of two)
    constexpr int MPW = 32 / LG;                                        // matrices per warp
    constexpr uint32_t P = 0x7FFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LG, r = lane % LG;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long mat = warp * MPW + sub; // matrix index = game * S + slice
    const long long game = mat / S;
    const int slice = (int)(mat - game * S);
    const bool live = game < B && r < S;
    uint32_t row[S];
#pragma unroll
    for (int c = 0; c < S; c++) {
        int v = 0;
        if (live) v = slab[game * G::GP + slice * G::RP + r * S + c];
        row[c] = v >= 0 ? (uint32_t)v : P - (uint32_t)(-v);
    }
    const uint32_t gmask = (LG == 32) ? 0xFFFFFFFFu : (((1u << LG) - 1u) << (sub * LG));
    bool used = !live;
    int rank = 0;
#pragma unroll
    for (int c = 0; c < S; c++) {
        const uint32_t cand = __ballot_sync(0xFFFFFFFFu, !used && row[c] != 0) & gmask;
        const bool has = cand != 0; // uniform within the group; all 32 lanes keep executing the same shuffles
        const int pl = has ? __ffs(cand) - 1 : lane;
        const uint32_t pv = __shfl_sync(0xFFFFFFFFu, row[c], pl);
        const uint32_t mine = row[c];
        const bool elim = has && !used && lane != pl && mine != 0;
        // columns <= c of the rows still in play are zero from here on: only k > c needs the update
        //   row[k] <- row[k] * pv - mine * pk  =  row[k] * pv + (P - mine) * pk   (mod P), one reduction for both products
        const uint32_t nmine = P - mine;
#pragma unroll
        for (int k = c + 1; k < S; k++) {
            const uint32_t pk = __shfl_sync(0xFFFFFFFFu, row[k], pl);
            if (elim) {
                const unsigned long long z = (unsigned long long)row[k] * pv + (unsigned long long)nmine * pk; // < 2^63
                unsigned long long r = (z & 0x7FFFFFFFull) + (z >> 31);                                        // < 2^33
                uint32_t q = (uint32_t)(r & 0x7FFFFFFFull) + (uint32_t)(r >> 31);
                row[k] = q >= P ? q - P : q;
            }
        }
        if (has && lane == pl) used = true;
        rank += has ? 1 : 0;
    }
    // every lane of a group counted the same pivots; add the slices of a game
    if (live && r == 0) atomicAdd(&ranks[game], rank);
}

// ------------------------------------------------------------------ K7
// 64-bit state key replacing the string key of utils.py:164-169 (dict key of
// the MCTS tree, act.py:37...210): a linear hash, sum over entries e of value * C_e,
// C_e = splitmix64(e+1) | 1, e the dense index (i*S+j)*S+k.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// key(T) = sum_e T[e] * C_e (mod 2^64), C_e = splitmix64(e + 1) | 1: one 64-bit multiply-add per non-zero entry, the
// constants in a shared-memory table built at kernel start; per-game sums through shared memory, one store per game.
template <int S>
__global__ void __launch_bounds__(256) state_key_kernel(const int8_t *__restrict__ slab, unsigned long long *__restrict__ keys,
                                                        long long B) {
    using G = Geo<S>;
    constexpr int TG = 256 / G::WR; // games per pass of the CTA
    __shared__ unsigned long long s_c[S * G::RP]; // C by slab offset (i, 4c+q); 0 in the row padding
    __shared__ unsigned long long s_key[TG];
    const int tid = threadIdx.x;
    for (int x = tid; x < S * G::RP; x += 256) {
        const int i = x / G::RP, jk = x % G::RP;
        s_c[x] = jk < G::S2 ? (splitmix64((unsigned long long)(i * G::S2 + jk + 1)) | 1ull) : 0ull;
    }
    const int gl = tid / G::WR, c = tid % G::WR;
    for (long long g0 = (long long)blockIdx.x * TG; g0 < B; g0 += (long long)gridDim.x * TG) {
        if (tid < TG) s_key[tid] = 0;
        __syncthreads();
        const long long g = g0 + gl;
        if (gl < TG && g < B) {
            unsigned long long h = 0;
            const uint32_t *col = reinterpret_cast<const uint32_t *>(slab + g * G::GP) + c;
#pragma unroll
            for (int i = 0; i < S; i++) {
                const uint32_t w = col[i * G::WR];
                if (w == 0) continue;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    h += (unsigned long long)(long long)(int8_t)((w >> (8 * q)) & 0xFFu) * s_c[i * G::RP + 4 * c + q];
            }
            if (h) atomicAdd(&s_key[gl], h);
        }
        __syncthreads();
        if (tid < TG && g0 + tid < B) keys[g0 + tid] = s_key[tid];
        __syncthreads();
    }
}

} // namespace tg

#define TG_SWITCH_S(S, ...) \
    switch (S) {            \
    case 4: { constexpr int kS = 4; __VA_ARGS__; } break;   \
    case 9: { constexpr int kS = 9; __VA_ARGS__; } break;   \
    case 16: { constexpr int kS = 16; __VA_ARGS__; } break; \
    default: return TG_E_ARG; \
    }

namespace tg {

template <int S, bool P16, bool STG>
static int launch_demo_sample_variant(unsigned grid, size_t smem, const uint8_t *tape, long long tape_step_stride, const int8_t *slab,
                                      long long N, int R, int dim_t, int replay_shift, const long long *idx, long long nb,
                                      float *states, float *scalars, long long *actions, float *rewards, cudaStream_t st) {
    auto kern = demo_sample_kernel<S, P16, STG>;
    if (STG) TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, STG ? smem : 0, st>>>(tape, tape_step_stride, slab, N, R, dim_t, replay_shift, idx, nb, states, scalars,
                                            actions, rewards);
    return TG_OK;
}

template <int S>
static int launch_demo_sample(const uint8_t *tape, long long tape_step_stride, const int8_t *slab, long long N, int R, int dim_t,
                              int replay_shift, const long long *idx, long long nb, float *states, float *scalars,
                              long long *actions, float *rewards, bool pack16, cudaStream_t st) {
    using G = Geo<S>;
    const unsigned grid = (unsigned)((nb * G::WR + 127) / 128);
    // staging needs 16-byte aligned records and room for the records of every sample a CTA touches
    const long long smem = ((128 + G::WR - 1) / G::WR + 1) * (long long)R * G::TP;
    const bool stage = (((uintptr_t)tape | (uintptr_t)tape_step_stride) & 15) == 0 && smem <= 64 * 1024;
#define TG_ARGS grid, (size_t)smem, tape, tape_step_stride, slab, N, R, dim_t, replay_shift, idx, nb, states, scalars, actions, rewards, st
    if (pack16) return stage ? launch_demo_sample_variant<S, true, true>(TG_ARGS) : launch_demo_sample_variant<S, true, false>(TG_ARGS);
    return stage ? launch_demo_sample_variant<S, false, true>(TG_ARGS) : launch_demo_sample_variant<S, false, false>(TG_ARGS);
#undef TG_ARGS
}

} // namespace tg

extern "C" {

int tg_demo_sample(const uint8_t *tape, int64_t tape_step_stride, const int8_t *slab, int64_t N, int R, int S, int dim_t,
                   int replay_shift, const int64_t *idx, int64_t nb, float *states, float *scalars, int64_t *actions,
                   float *rewards, void *stream) {
    if (!tg::supported_S(S) || N < 0 || R < 1 || dim_t < 1 || nb < 0) return TG_E_ARG;
    if (nb == 0) return TG_OK;
    if (!tape || !slab || !idx || !states || !scalars || !actions || !rewards) return TG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // tokens are <= 2 * 4 (tape contract), so |coefficient| <= cmax; the 16-bit packed head is exact while
    // 127 + R * cmax^3 fits an int16 lane
    const long long cmax = replay_shift > 8 - replay_shift ? replay_shift : 8 - replay_shift;
    const bool pack16 = replay_shift >= 0 && replay_shift <= 8 && 127 + (long long)R * cmax * cmax * cmax <= 32767;
    TG_SWITCH_S(S, {
        const int rc = tg::launch_demo_sample<kS>(tape, tape_step_stride, slab, N, R, dim_t, replay_shift, (const long long *)idx, nb,
                                                  states, scalars, (long long *)actions, rewards, pack16, st);
        if (rc != TG_OK) return rc;
    });
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int tg_slice_rank(const int8_t *slab, int32_t *ranks, int64_t B, int S, void *stream) {
    if (!tg::supported_S(S) || B < 0) return TG_E_ARG;
    if (B == 0) return TG_OK;
    if (!slab || !ranks) return TG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    TG_CUDA(cudaMemsetAsync(ranks, 0, (size_t)B * 4, st));
    TG_SWITCH_S(S, {
        constexpr int LG = kS <= 4 ? 4 : (kS <= 8 ? 8 : (kS <= 16 ? 16 : 32));
        const long long warps = (B * kS + (32 / LG) - 1) / (32 / LG);
        tg::slice_rank_kernel<kS><<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(slab, ranks, B);
    });
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int tg_state_key(const int8_t *slab, uint64_t *keys, int64_t B, int S, void *stream) {
    if (!tg::supported_S(S) || B < 0) return TG_E_ARG;
    if (B == 0) return TG_OK;
    if (!slab || !keys) return TG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    TG_SWITCH_S(S, {
        const long long tiles = (B + (256 / tg::Geo<kS>::WR) - 1) / (256 / tg::Geo<kS>::WR);
        tg::state_key_kernel<kS><<<(unsigned)(tiles < 148 * 16 ? tiles : 148 * 16), 256, 0, st>>>(slab, (unsigned long long *)keys, B);
    });
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

} // extern "C"
