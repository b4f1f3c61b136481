// tg_expand.cu -- K8: batched leaf expansion (C ABI: tg_expand_children).
//
// Reference restated: get_child_states (act.py:266-275) -- for each of the k sampled actions of a state,
// new_head = head - u (x) v (x) w -- together with what extend_tree does with the children right away:
// remove_null_actions (utils.py:191-194), tensor_factorized on the head (utils.py:181-188, act.py:177) and the
// "already in the tree" lookup by state key (act.py:185-195; key = tg_state_key instead of utils.py:164-169 strings).
// The reference does this for one state at a time; here B states expand at once:
//     children[b][c] = parents[b] - rank1(tape[b][c]),  c < k,
// with per-child flags (terminal / null action / range), non-zero count and 64-bit state key.
//
// The parent is read ONCE: thread (parent, word column) keeps its S row words in registers and produces the k
// children one after the other into a two-stage shared-memory tile; each child game leaves with a TMA bulk store
// while the next one is computed.  HBM traffic per child: GP * (1 + 1/k) + TP + 13 bytes.
// The state key is linear (sum_e T[e] * C_e mod 2^64, tg_state_key), so a child's key is the parent's key (hashed
// once per parent) minus sum_e (u_i v_j w_k) * C_e over the few entries its rank-1 action changes -- not a second
// pass over the child.  The constants C_e of a thread's word column live in a shared-memory table.
#include "tg_step.cuh"

namespace tg {

template <int S, int NT>
struct ExpCfg {
    using G = Geo<S>;
    static constexpr int TG = NT / G::WR; // parents per CTA (one word column per thread)
    static constexpr int ACTIVE = TG * G::WR;
    static constexpr int STAGE_BYTES = TG * G::GP;
    static __host__ __device__ constexpr int tok_bytes(int k) { return (TG * k * G::TP + 15) & ~15; }
    // tokens, 2 child stages, per-(stage, game) partial word and key
    static constexpr int CTAB_BYTES = S * G::RP * 8; // key constants by slab offset
    static __host__ __device__ constexpr int smem_bytes(int k) {
        return tok_bytes(k) + 2 * STAGE_BYTES + 2 * TG * 4 + 3 * TG * 8 + 16 + CTAB_BYTES;
    }
};

__device__ __forceinline__ unsigned long long splitmix64_e(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int S, int NT, bool KEYS>
__global__ void __launch_bounds__(NT)
    expand_kernel(const int8_t *__restrict__ parents, const uint8_t *__restrict__ tape, int k, int8_t *__restrict__ children,
                  uint8_t *__restrict__ flags, int32_t *__restrict__ nnz, unsigned long long *__restrict__ keys, long long B,
                  int shift) {
    using C = ExpCfg<S, NT>;
    using G = Geo<S>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_tok = smem;                                                            // [TG][k][TP]
    uint8_t *s_out = smem + C::tok_bytes(k);                                          // [2][TG][GP]
    uint32_t *s_part = reinterpret_cast<uint32_t *>(s_out + 2 * C::STAGE_BYTES);      // [2][TG]
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_part + 2 * C::TG); // [2][TG] key deltas, [TG] parent keys
    unsigned long long *s_pkey = s_key + 2 * C::TG;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_pkey + C::TG);
    unsigned long long *s_c = reinterpret_cast<unsigned long long *>(s_bar + 2); // [S][RP]

    const int tid = threadIdx.x;
    const long long g0 = (long long)blockIdx.x * C::TG;
    const int ng = (int)min((long long)C::TG, B - g0);
    const int g = tid / G::WR;
    const bool active = tid < C::ACTIVE && g < ng;
    Lane<S> L;
    L.init(tid < C::ACTIVE ? tid % G::WR : 0);

    if (tid == 0) {
        mbar_init(s_bar, 1);
        mbar_fence_init();
    }
    // game padding of both child stages is zero and stays zero (threads only write their word columns of the S rows)
    for (int w = tid; w < 2 * C::STAGE_BYTES / 4; w += NT) reinterpret_cast<uint32_t *>(s_out)[w] = 0;
    for (int i = tid; i < 2 * C::TG; i += NT) s_part[i] = 0, s_key[i] = 0;
    for (int i = tid; i < C::TG; i += NT) s_pkey[i] = 0;
    if constexpr (KEYS)
        for (int x = tid; x < S * G::RP; x += NT) {
            const int i = x / G::RP, jk = x % G::RP;
            s_c[x] = jk < G::S2 ? (splitmix64_e((unsigned long long)(i * G::S2 + jk + 1)) | 1ull) : 0ull;
        }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(s_bar, (uint32_t)(ng * k * G::TP));
        bulk_g2s(s_tok, tape + g0 * k * G::TP, (uint32_t)(ng * k * G::TP), s_bar);
    }
    // the parent's rows of this thread's word column, offset-binary
    uint32_t row[S];
    if (active) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(parents + (g0 + g) * G::GP) + L.c;
#pragma unroll
        for (int i = 0; i < S; i++) row[i] = src[i * G::WR] ^ H4;
    } else {
#pragma unroll
        for (int i = 0; i < S; i++) row[i] = H4;
    }
    // sum_q (signed byte q of word) * C[row i][4c + q]
    auto word_key = [&](int i, uint32_t word) {
        unsigned long long h = 0;
#pragma unroll
        for (int q = 0; q < 4; q++)
            h += (unsigned long long)(long long)(int8_t)((word >> (8 * q)) & 0xFFu) * s_c[i * G::RP + 4 * L.c + q];
        return h;
    };
    if constexpr (KEYS) {
        if (active) { // the parent's key, once
            unsigned long long h = 0;
#pragma unroll
            for (int i = 0; i < S; i++) {
                const uint32_t t = row[i] ^ H4;
                if (t != 0) h += word_key(i, t);
            }
            if (h) atomicAdd(&s_pkey[g], h);
        }
    }
    mbar_wait(s_bar, 0);

    for (int c = 0; c < k; c++) {
        const int st = c & 1;
        uint8_t *stage = s_out + st * C::STAGE_BYTES;
        // the bulk stores of child c-2 have finished reading this stage (issuing threads waited, then the barrier below)
        if (active) {
            const uint8_t *tok = s_tok + ((size_t)g * k + c) * G::TP;
            const int32_t vw = pack_vw<S>(tok, L, shift);
            const uint4 ut = *reinterpret_cast<const uint4 *>(tok);
            const uint32_t uw[4] = {ut.x, ut.y, ut.z, ut.w};
            const uint32_t nvw = (uint32_t)(-vw);
            const uint32_t vwb = ((uint32_t)(-vw) + H4) ^ H4; // -(v w) of the four entries as two's complement bytes
            uint32_t cnt = 0, rng = 0;
            int uany = 0;
            unsigned long long h = 0;
            uint32_t *col = reinterpret_cast<uint32_t *>(stage + (size_t)g * G::GP) + L.c;
#pragma unroll
            for (int i = 0; i < S; i++) {
                const int negu = coef_u(uw, i, shift);
                const uint32_t t = (row[i] + (uint32_t)negu * nvw) ^ H4; // child word, two's complement bytes
                col[i * G::WR] = t;
                cnt += ((((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & L.hv) >> 7;
                rng |= t ^ (t << 1);
                uany |= negu;
                if constexpr (KEYS) {
                    // the word changes by -u_i * (v w bytes): its key contribution changes by -u_i * word_key(vw bytes)
                    if (negu != 0 && vw != 0) h += (unsigned long long)(long long)negu * word_key(i, vwb);
                }
            }
            atomicAdd(&s_part[st * C::TG + g], make_partial(byte_sum(cnt), vw != 0 && uany != 0,
                                                            (rng & L.hv) != 0 || tokens_out_of_range<S>(tok, L, shift)));
            if constexpr (KEYS)
                if (h) atomicAdd(&s_key[st * C::TG + g], h);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid < ng) {
            const long long child = (g0 + tid) * k + c;
            bulk_s2g(children + child * G::GP, stage + (size_t)tid * G::GP, (uint32_t)G::GP);
            bulk_commit();
            const uint32_t sum = s_part[st * C::TG + tid];
            flags[child] = (uint8_t)partial_flags(sum);
            nnz[child] = (int32_t)(sum & 0xFFFFu);
            s_part[st * C::TG + tid] = 0;
            if constexpr (KEYS) {
                keys[child] = s_pkey[tid] + s_key[st * C::TG + tid];
                s_key[st * C::TG + tid] = 0;
            }
            bulk_wait_read<1>(); // the store of child c-1 (other stage) has drained: safe to overwrite next iteration
        }
        // (the barrier at the end of the NEXT iteration's compute orders "stage drained" before its reuse at c+2;
        //  the one here orders the reset of s_part / s_key before the next child's atomics)
        __syncthreads();
    }
    if (tid < ng) bulk_wait<0>();
}

template <int S, int NT>
static int launch_expand(const int8_t *parents, const uint8_t *tape, int k, int8_t *children, uint8_t *flags, int32_t *nnz,
                         unsigned long long *keys, long long B, int shift, cudaStream_t st) {
    using C = ExpCfg<S, NT>;
    const int smem = C::smem_bytes(k);
    if (smem > 200 * 1024) return TG_E_ARG;
    const long long grid = (B + C::TG - 1) / C::TG;
    if (grid > 0x7FFFFFFFLL) return TG_E_ARG;
    if (keys) {
        auto kern = expand_kernel<S, NT, true>;
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<(int)grid, NT, smem, st>>>(parents, tape, k, children, flags, nnz, keys, B, shift);
    } else {
        auto kern = expand_kernel<S, NT, false>;
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<(int)grid, NT, smem, st>>>(parents, tape, k, children, flags, nnz, keys, B, shift);
    }
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

} // namespace tg

extern "C" int tg_expand_children(const int8_t *parents, const uint8_t *tape, int k, int8_t *children, uint8_t *flags,
                                  int32_t *nnz, uint64_t *keys, int64_t B, int S, int shift, void *stream) {
    if (!tg::supported_S(S) || B < 0 || k < 1 || shift < 1 || shift > 4) return TG_E_ARG;
    if (B == 0) return TG_OK;
    if (!parents || !tape || !children || !flags || !nnz) return TG_E_ARG;
    if (((uintptr_t)parents | (uintptr_t)tape | (uintptr_t)children) & 15) return TG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (S) {
    case 4: return tg::launch_expand<4, 256>(parents, tape, k, children, flags, nnz, (unsigned long long *)keys, B, shift, st);
    case 9: return tg::launch_expand<9, 256>(parents, tape, k, children, flags, nnz, (unsigned long long *)keys, B, shift, st);
    case 16: return tg::launch_expand<16, 256>(parents, tape, k, children, flags, nnz, (unsigned long long *)keys, B, shift, st);
    }
    return TG_E_ARG;
}
