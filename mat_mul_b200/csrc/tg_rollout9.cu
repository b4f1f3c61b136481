// tg_rollout9.cu -- K2r: the fused K-step rollout (tg_rollout / tg_replay) at 9x9x9, one thread per ROW of a game.
//
// Reference restated: SyntheticDemoDataset._take_actions (datasets.py:144-153) and the greedy loop of
// training.py:336-342 around _take_action: K successive transitions T <- T - u (x) v (x) w, the game frozen once its
// head is all zero (the break at act.py:49); same contract, flags and outputs as the word-column kernel of tg_rollout.cu.
//
// In tg_rollout.cu a thread owns one word COLUMN of a game (21 threads per game, 9 row words each): every thread extracts
// all nine u coefficients of every step and builds its own pack(v w) word, whose four entries straddle two v_j -- 77 warp-
// instructions per game-step, issue slots 80 % busy, 0.17 of the HBM roofline.  Here a thread owns ROW i of a game, nine
// lanes per game, three games per warp, and keeps the row in registers as NINE RUNS of three words -- run j = entries
// (i, j, 0..8) as 4 + 4 + 1 packed bytes (offset-binary as in tg_step.cuh; the 9-byte runs of the slab are unpacked once when
// the game comes in and packed once when it leaves).  In that form a step needs no pack(v w) at all:
//     run[j][m] += (-u_i v_j) * W_m ,   W_m = the three words of pack(w) in integer form,
// i.e. one u coefficient (a byte load), nine v coefficients (three packed subtractions, nine sign-extending PRMTs), three W
// words (two funnel shifts), nine products and 27 independent IMADs per lane -- no exchange between lanes, no shared-
// memory traffic besides the token record, and "is the game solved" is a checksum over nine of the 27 words plus one ballot (confirmed word by word when all nine rows pass) (the warp
// owns the whole game: no shared-memory votes, no atomics, no CTA-level barrier).  Tokens stream through the same TMA ring
// as in tg_rollout.cu; start states come in and results leave with one bulk copy per game through a per-warp stage.
#include "tg_rows.cuh"
#include "tg_step.cuh"

namespace tg {
namespace roll9 {

constexpr int NW = 8;  // compute warps per CTA
constexpr int NST = 4; // token ring depth
// 16x16x16 runs the same kernel: rows of 16 aligned runs of four words (no padding, no unpacking), 16 lanes per game, two
// games per warp.
template <int S_>
struct RowGeo {
    using G = Geo<S_>;
    static constexpr int S = S_, RP = G::RP, GP = G::GP, TP = G::TP, WR = G::RP / 4;
    static constexpr int KW = (S + 3) / 4;          // words per run
    static constexpr int GPW = 32 / S, TG = NW * GPW; // games per warp, games per CTA
    static constexpr int TOK_BYTES = TG * TP;       // one ring stage: the step's records of the CTA's games
    // state stage of a warp: 9x9x9 games as they lie in the slab (rows 21 words apart: the nine lanes of a game hit nine
    // banks); 16x16x16 rows are 64 words = one bank for all sixteen lanes, so every row is a bulk copy of its own to a
    // pitch of 68 words (2-way instead of 16-way conflicts on the way in and out)
    static constexpr bool ROWCOPY = RP % 128 == 0;
    static constexpr int ROWPITCH = ROWCOPY ? RP + 16 : RP;
    static constexpr int GPITCH = (ROWCOPY ? S * ROWPITCH : GP) + 16; // game pitch (banks of a warp's games 4 words apart)
    static constexpr int STAGE_BYTES = GPW * GPITCH;
    static constexpr int WARP_BYTES = STAGE_BYTES + 32 * 4; // + one word per lane for the final reductions
    static constexpr int SMEM_BYTES = NST * TOK_BYTES + NW * WARP_BYTES + (2 * NST + NW) * 8;
    static constexpr uint32_t GROUP = (1u << S) - 1u; // the lanes of one game in a ballot
    static_assert(TOK_BYTES % 16 == 0 && WARP_BYTES % 16 == 0, "bulk-copy alignment");
    static_assert(S == 9 || S == 16, "row-owner rollout: 9x9x9 and 16x16x16");
};

template <int B>
__device__ __forceinline__ int sx(uint32_t w) { // sign-extended byte B
    constexpr uint32_t sel = (uint32_t)B | ((uint32_t)(B | 8) << 4) | ((uint32_t)(B | 8) << 8) | ((uint32_t)(B | 8) << 12);
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0u), "r"(sel));
    return (int)d;
}

template <int S, bool FREEZE>
__global__ void __launch_bounds__(32 * (NW + 1), S == 9 ? 3 : 2)
    rollout_rows_kernel(const int8_t *__restrict__ slab_in, const uint8_t *__restrict__ tape, long long tape_step_stride, int K,
                         int8_t *__restrict__ slab_out, uint8_t *__restrict__ flags, int32_t *__restrict__ nnz,
                         int32_t *__restrict__ steps, long long B, int shift, int chk) {
    using R = RowGeo<S>;
    constexpr int RP = R::RP, GP = R::GP, TP = R::TP, WR = R::WR, KW = R::KW, GPW = R::GPW, TG = R::TG, TOK_BYTES = R::TOK_BYTES,
                  GPITCH = R::GPITCH, STAGE_BYTES = R::STAGE_BYTES, WARP_BYTES = R::WARP_BYTES, ROWPITCH = R::ROWPITCH;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_tok = smem;                                                     // [NST][TG][TP]
    uint8_t *s_warp = smem + NST * TOK_BYTES;                                  // [NW][WARP_BYTES]
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_warp + NW * WARP_BYTES); // [NST] tokens of a step have landed
    uint64_t *s_empty = s_full + NST;                                          // [NST] ... every compute warp has read them
    uint64_t *s_in = s_empty + NST;                                            // [NW]  the warp's start states have landed

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long g0 = (long long)blockIdx.x * TG;
    const int ng = (int)min((long long)TG, B - g0);
    if (tid == 0) {
        for (int s = 0; s < NST; s++) mbar_init(&s_full[s], 1), mbar_init(&s_empty[s], NW);
        for (int w = 0; w < NW; w++) mbar_init(&s_in[w], 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NW) {
        // ---------------- producer warp: one lane streams the tokens of step t into ring stage t % NST
        if (lane == 0) {
            const uint8_t *src = tape + g0 * TP;
            for (int t = 0; t < K; t++, src += tape_step_stride) {
                const int st = t & (NST - 1);
                if (t >= NST) mbar_wait(&s_empty[st], (uint32_t)(t / NST - 1) & 1u);
                mbar_expect_tx(&s_full[st], (uint32_t)(ng * TP));
                bulk_g2s(s_tok + st * TOK_BYTES, src, (uint32_t)(ng * TP), &s_full[st]);
            }
        }
        return;
    }

    // ---------------- compute warps
    uint8_t *s_stage = s_warp + warp * WARP_BYTES;                          // [GPW][GPITCH] start states, later the results
    uint32_t *s_red = reinterpret_cast<uint32_t *>(s_stage + STAGE_BYTES);  // [32]
    const int wg0 = warp * GPW;                                             // first game of the warp inside the CTA tile
    const int nwg = max(0, min(GPW, ng - wg0));                             // games this warp really has
    const int q = lane / S, i = lane - q * S; // this lane's game and row (9x9x9: lanes 27..31 have q == 3 and idle)
    const bool owner = q < nwg;
    if (lane == 0 && nwg > 0) mbar_expect_tx(&s_in[warp], (uint32_t)(nwg * GP));
    __syncwarp();
    if constexpr (R::ROWCOPY) {
        if (owner) bulk_g2s(s_stage + q * GPITCH + i * ROWPITCH, slab_in + (g0 + wg0 + q) * GP + i * RP, RP, &s_in[warp]);
    } else if (lane == 0) {
        for (int qq = 0; qq < nwg; qq++) bulk_g2s(s_stage + qq * GPITCH, slab_in + (g0 + wg0 + qq) * GP, GP, &s_in[warp]);
    }

    uint32_t run[S][KW]; // run j = entries (i, j, 0..S-1) as packed words, offset-binary; 9x9x9: bytes 1..3 of word 2 are padding (0x80)
    uint32_t bad = 0;
    bool alive = owner;
    if (nwg > 0) mbar_wait(&s_in[warp], 0);
    {
        uint32_t r[WR + 2];
        uint32_t nzw = 0;
        if (owner) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(s_stage + q * GPITCH + i * ROWPITCH);
#pragma unroll
            for (int w = 0; w < WR; w++) r[w] = src[w];
        } else {
#pragma unroll
            for (int w = 0; w < WR; w++) r[w] = 0;
        }
        unpack_row<S>(r, run);
#pragma unroll
        for (int j = 0; j < S; j++)
#pragma unroll
            for (int m = 0; m < KW; m++) {
                nzw |= run[j][m];
                run[j][m] ^= H4;
                bad |= ~(run[j][m] ^ (run[j][m] << 1)); // the start state must already be inside [-64,63]
            }
        s_red[lane] = nzw;
    }
    __syncwarp();
    if (FREEZE && owner) { // a game that starts at the zero tensor is frozen from the start (0 steps)
        uint32_t any = 0;
#pragma unroll
        for (int rr = 0; rr < S; rr++) any |= s_red[q * S + rr];
        alive = any != 0;
    }
    __syncwarp();
    int until = chk, my_steps = 0;
    constexpr uint32_t ZSUM = (uint32_t)((unsigned long long)S * H4); // checksum (first word of every run) of an all-zero row
    const uint32_t sh4 = (uint32_t)shift * ONES4;
    const uint8_t *tok_mine = s_tok + (wg0 + (owner ? q : 0)) * TP;

    auto step = [&](int t, int st, uint32_t parity) {
        mbar_wait(&s_full[st], parity);
        bool zero_row = true;
        if (alive) {
            const uint8_t *tok = tok_mine + st * TOK_BYTES;
            const int nu = shift - (int)tok[i]; // -u_i
            int c[S], W[KW];                    // -u_i v_j; pack(w) in integer form
            if constexpr (S == 9) { // v = bytes 9..17, w = bytes 18..26 of the record
                const uint4 qa = *reinterpret_cast<const uint4 *>(tok), qb = *reinterpret_cast<const uint4 *>(tok + 16);
                const uint32_t cv2 = ((qa.z | H4) - sh4) ^ H4, cv3 = ((qa.w | H4) - sh4) ^ H4, cv4 = ((qb.x | H4) - sh4) ^ H4;
                c[0] = nu * sx<1>(cv2), c[1] = nu * sx<2>(cv2), c[2] = nu * sx<3>(cv2), c[3] = nu * sx<0>(cv3), c[4] = nu * sx<1>(cv3);
                c[5] = nu * sx<2>(cv3), c[6] = nu * sx<3>(cv3), c[7] = nu * sx<0>(cv4), c[8] = nu * sx<1>(cv4);
                W[0] = (int)(__funnelshift_r(qb.x, qb.y, 16) - sh4), W[1] = (int)(__funnelshift_r(qb.y, qb.z, 16) - sh4);
                W[2] = (int)((qb.z >> 16) & 0xFFu) - shift;
            } else { // v = bytes 16..31, w = bytes 32..47
                const uint4 qv = *reinterpret_cast<const uint4 *>(tok + 16), qw = *reinterpret_cast<const uint4 *>(tok + 32);
                const uint32_t vw[4] = {qv.x, qv.y, qv.z, qv.w}, ww[4] = {qw.x, qw.y, qw.z, qw.w};
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const uint32_t cv = ((vw[m] | H4) - sh4) ^ H4;
                    c[4 * m] = nu * sx<0>(cv), c[4 * m + 1] = nu * sx<1>(cv), c[4 * m + 2] = nu * sx<2>(cv), c[4 * m + 3] = nu * sx<3>(cv);
                    W[m] = (int)(ww[m] - sh4);
                }
            }
            uint32_t sum = 0;
#pragma unroll
            for (int j = 0; j < S; j++) {
#pragma unroll
                for (int m = 0; m < KW; m++) run[j][m] += (uint32_t)(c[j] * W[m]);
                sum += run[j][0]; // a cheap necessary condition for a zero row: the first words of its runs
            }
            if (--until == 0) { // all entries still in [-64,63]? then chk more steps cannot alias the packed form
                until = chk;
#pragma unroll
                for (int j = 0; j < S; j++)
#pragma unroll
                    for (int m = 0; m < KW; m++) bad |= ~(run[j][m] ^ (run[j][m] << 1));
            }
            if (i < TP / 4) { // tape contract (tg_step.cuh): every token <= 2 * shift, else the packed update may have aliased
                const uint32_t x = reinterpret_cast<const uint32_t *>(tok)[i];
                if (((((x & 0x7F7F7F7Fu) + (uint32_t)(0x7F - 2 * shift) * ONES4) | x) & H4) != 0) bad = 0xFFFFFFFFu;
            }
            my_steps = t + 1;
            if (FREEZE) zero_row = sum == ZSUM; // necessary for a zero row; confirmed below when the whole game passes
        }
        if (FREEZE) {
            uint32_t m = __ballot_sync(0xFFFFFFFFu, zero_row);
            bool solved = alive && ((m >> (S * q)) & R::GROUP) == R::GROUP;
            if (__any_sync(0xFFFFFFFFu, solved)) { // rare: every row of a game has the zero checksum -- compare word by word
                bool exact = true;
#pragma unroll
                for (int j = 0; j < S; j++)
#pragma unroll
                    for (int mm = 0; mm < KW; mm++) exact = exact && run[j][mm] == H4;
                m = __ballot_sync(0xFFFFFFFFu, !alive || exact);
                solved = alive && ((m >> (S * q)) & R::GROUP) == R::GROUP;
            }
            if (solved) alive = false; // solved by this step: frozen at the zero tensor
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_empty[st])) : "memory");
    };
    int t = 0;
    for (; t + NST <= K; t += NST) {
        const uint32_t parity = (uint32_t)(t / NST) & 1u;
#pragma unroll
        for (int s = 0; s < NST; s++) step(t + s, s, parity);
    }
    for (; t < K; t++) step(t, t & (NST - 1), (uint32_t)(t / NST) & 1u);

    // ---------------- results: the runs packed back into the row's 21 words, per-game nnz / flags / steps
    uint32_t cnt = 0;
    if (owner) {
#pragma unroll
        for (int j = 0; j < S; j++)
#pragma unroll
            for (int m = 0; m < KW; m++) {
                bad |= ~(run[j][m] ^ (run[j][m] << 1));
                run[j][m] ^= H4; // two's complement bytes (padding bytes of word 2: zero)
                cnt += (uint32_t)__popc(nonzero_mask(run[j][m]));
            }
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_stage + q * GPITCH + i * ROWPITCH);
        uint32_t words[WR];
        pack_row<S>(run, words);
#pragma unroll
        for (int w = 0; w < WR; w++) dst[w] = words[w];
        if constexpr (GP > S * RP) {
            if (i == 0) { // game padding (9x9x9: bytes 756..767) stays zero
                uint32_t *pad = reinterpret_cast<uint32_t *>(s_stage + q * GPITCH + S * RP);
#pragma unroll
                for (int x = 0; x < (GP - S * RP) / 4; x++) pad[x] = 0u;
            }
        }
    }
    s_red[lane] = cnt;
    const uint32_t badm = __ballot_sync(0xFFFFFFFFu, (bad & H4) != 0);
    fence_proxy_async();
    __syncwarp();
    if (owner && i == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int rr = 0; rr < S; rr++) total += s_red[q * S + rr];
        const long long gidx = g0 + wg0 + q;
        flags[gidx] = (uint8_t)((total == 0 ? TG_FLAG_TERMINAL : 0u) | (((badm >> (S * q)) & R::GROUP) != 0 ? TG_FLAG_RANGE : 0u));
        nnz[gidx] = (int32_t)total;
        if (steps) steps[gidx] = my_steps;
    }
    if constexpr (R::ROWCOPY) {
        if (owner) {
            bulk_s2g(slab_out + (g0 + wg0 + q) * GP + i * RP, s_stage + q * GPITCH + i * ROWPITCH, RP);
            bulk_commit();
            bulk_wait_read<0>();
        }
    } else if (lane == 0 && nwg > 0) {
        for (int qq = 0; qq < nwg; qq++) bulk_s2g(slab_out + (g0 + wg0 + qq) * GP, s_stage + qq * GPITCH, GP);
        bulk_commit();
        bulk_wait_read<0>();
    }
}

} // namespace roll9

template <int S>
static int launch_rows(const int8_t *slab_in, const uint8_t *tape, long long stride, int K, int8_t *slab_out, uint8_t *flags,
                       int32_t *nnz, int32_t *steps, long long B, int shift, int freeze, cudaStream_t st) {
    using R = roll9::RowGeo<S>;
    const long long grid = (B + R::TG - 1) / R::TG;
    if (grid > 0x7FFFFFFFLL) return TG_E_ARG;
    const int s3 = shift * shift * shift;
    const int chk = s3 >= 64 ? 1 : 64 / s3;
    if (freeze) {
        auto kern = roll9::rollout_rows_kernel<S, true>;
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, R::SMEM_BYTES));
        kern<<<(int)grid, 32 * (roll9::NW + 1), R::SMEM_BYTES, st>>>(slab_in, tape, stride, K, slab_out, flags, nnz, steps, B, shift, chk);
    } else {
        auto kern = roll9::rollout_rows_kernel<S, false>;
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, R::SMEM_BYTES));
        kern<<<(int)grid, 32 * (roll9::NW + 1), R::SMEM_BYTES, st>>>(slab_in, tape, stride, K, slab_out, flags, nnz, steps, B, shift, chk);
    }
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int launch_rollout_rows(const int8_t *slab_in, const uint8_t *tape, long long stride, int K, int8_t *slab_out, uint8_t *flags,
                        int32_t *nnz, int32_t *steps, long long B, int S, int shift, int freeze, cudaStream_t st) {
    if (S == 9) return launch_rows<9>(slab_in, tape, stride, K, slab_out, flags, nnz, steps, B, shift, freeze, st);
    if (S == 16) return launch_rows<16>(slab_in, tape, stride, K, slab_out, flags, nnz, steps, B, shift, freeze, st);
    return TG_E_ARG;
}

} // namespace tg
