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
#include "tg_step.cuh"

namespace tg {
namespace roll9 {

constexpr int S = 9, RP = 84, GP = 768, TP = 32, WR = 21;
constexpr int NW = 8, GPW = 3, TG = NW * GPW; // compute warps per CTA, games per warp, games per CTA
constexpr int NST = 4;                        // token ring depth
constexpr int TOK_BYTES = TG * TP;            // one ring stage: the step's records of the CTA's games
constexpr int GPITCH = GP + 16;               // game pitch of the state stage (banks of the three games 4 words apart)
constexpr int STAGE_BYTES = GPW * GPITCH;
constexpr int WARP_BYTES = STAGE_BYTES + 32 * 4; // + one word per lane for the final reductions
constexpr int SMEM_BYTES = NST * TOK_BYTES + NW * WARP_BYTES + (2 * NST + NW) * 8;
static_assert(TOK_BYTES % 16 == 0 && WARP_BYTES % 16 == 0, "bulk-copy alignment");

template <int B>
__device__ __forceinline__ int sx(uint32_t w) { // sign-extended byte B
    constexpr uint32_t sel = (uint32_t)B | ((uint32_t)(B | 8) << 4) | ((uint32_t)(B | 8) << 8) | ((uint32_t)(B | 8) << 12);
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0u), "r"(sel));
    return (int)d;
}

template <bool FREEZE>
__global__ void __launch_bounds__(32 * (NW + 1), 3)
    rollout_rows9_kernel(const int8_t *__restrict__ slab_in, const uint8_t *__restrict__ tape, long long tape_step_stride, int K,
                         int8_t *__restrict__ slab_out, uint8_t *__restrict__ flags, int32_t *__restrict__ nnz,
                         int32_t *__restrict__ steps, long long B, int shift, int chk) {
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
    if (lane == 0 && nwg > 0) {
        mbar_expect_tx(&s_in[warp], (uint32_t)(nwg * GP));
        for (int q = 0; q < nwg; q++) bulk_g2s(s_stage + q * GPITCH, slab_in + (g0 + wg0 + q) * GP, GP, &s_in[warp]);
    }
    const int q = lane / S, i = lane - q * S; // this lane's game and row (lanes 27..31: q == 3, idle)
    const bool owner = q < nwg;

    uint32_t run[S][3]; // run j = entries (i, j, 0..3 | 4..7 | 8), offset-binary; bytes 1..3 of word 2 are padding (0x80)
    uint32_t bad = 0;
    bool alive = owner;
    if (nwg > 0) mbar_wait(&s_in[warp], 0);
    {
        uint32_t r[WR + 1];
        uint32_t nzw = 0;
        if (owner) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(s_stage + q * GPITCH + i * RP);
#pragma unroll
            for (int w = 0; w < WR; w++) r[w] = src[w];
            r[WR - 1] &= 0x000000FFu; // entries 81..83 of a row are padding
        } else {
#pragma unroll
            for (int w = 0; w < WR; w++) r[w] = 0;
        }
        r[WR] = 0;
#pragma unroll
        for (int j = 0; j < S; j++) {
            const int o = 9 * j, w0 = o >> 2, sh = 8 * (o & 3);
            const uint32_t x0 = __funnelshift_r(r[w0], r[w0 + 1], sh), x1 = __funnelshift_r(r[w0 + 1], r[w0 + 2 <= WR ? w0 + 2 : WR], sh);
            const int o8 = o + 8;
            const uint32_t x2 = (r[o8 >> 2] >> (8 * (o8 & 3))) & 0xFFu;
            nzw |= x0 | x1 | x2;
            run[j][0] = x0 ^ H4, run[j][1] = x1 ^ H4, run[j][2] = x2 ^ H4;
#pragma unroll
            for (int m = 0; m < 3; m++) bad |= ~(run[j][m] ^ (run[j][m] << 1)); // the start state must already be inside [-64,63]
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
    constexpr uint32_t ZSUM = (uint32_t)(9ull * H4); // checksum (first word of every run) of an all-zero row
    const uint32_t sh4 = (uint32_t)shift * ONES4;
    const uint8_t *tok_mine = s_tok + (wg0 + (owner ? q : 0)) * TP;

    auto step = [&](int t, int st, uint32_t parity) {
        mbar_wait(&s_full[st], parity);
        bool zero_row = true;
        if (alive) {
            const uint8_t *tok = tok_mine + st * TOK_BYTES;
            const uint4 qa = *reinterpret_cast<const uint4 *>(tok), qb = *reinterpret_cast<const uint4 *>(tok + 16);
            const int nu = shift - (int)tok[i]; // -u_i
            // v = bytes 9..17, w = bytes 18..26 of the record
            const uint32_t cv2 = ((qa.z | H4) - sh4) ^ H4, cv3 = ((qa.w | H4) - sh4) ^ H4, cv4 = ((qb.x | H4) - sh4) ^ H4;
            const int c[S] = {nu * sx<1>(cv2), nu * sx<2>(cv2), nu * sx<3>(cv2), nu * sx<0>(cv3), nu * sx<1>(cv3),
                              nu * sx<2>(cv3), nu * sx<3>(cv3), nu * sx<0>(cv4), nu * sx<1>(cv4)};
            const int W0 = (int)(__funnelshift_r(qb.x, qb.y, 16) - sh4), W1 = (int)(__funnelshift_r(qb.y, qb.z, 16) - sh4);
            const int W2 = (int)((qb.z >> 16) & 0xFFu) - shift;
            uint32_t sum = 0;
#pragma unroll
            for (int j = 0; j < S; j++) {
                run[j][0] += (uint32_t)(c[j] * W0), run[j][1] += (uint32_t)(c[j] * W1), run[j][2] += (uint32_t)(c[j] * W2);
                sum += run[j][0]; // a cheap necessary condition for a zero row: the first words of its nine runs
            }
            if (--until == 0) { // all entries still in [-64,63]? then chk more steps cannot alias the packed form
                until = chk;
#pragma unroll
                for (int j = 0; j < S; j++)
#pragma unroll
                    for (int m = 0; m < 3; m++) bad |= ~(run[j][m] ^ (run[j][m] << 1));
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
            bool solved = alive && ((m >> (S * q)) & 0x1FFu) == 0x1FFu;
            if (__any_sync(0xFFFFFFFFu, solved)) { // rare: every row of a game has the zero checksum -- compare word by word
                bool exact = true;
#pragma unroll
                for (int j = 0; j < S; j++)
#pragma unroll
                    for (int mm = 0; mm < 3; mm++) exact = exact && run[j][mm] == H4;
                m = __ballot_sync(0xFFFFFFFFu, !alive || exact);
                solved = alive && ((m >> (S * q)) & 0x1FFu) == 0x1FFu;
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
            for (int m = 0; m < 3; m++) {
                bad |= ~(run[j][m] ^ (run[j][m] << 1));
                run[j][m] ^= H4; // two's complement bytes (padding bytes of word 2: zero)
                cnt += (uint32_t)__popc(nonzero_mask(run[j][m]));
            }
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_stage + q * GPITCH + i * RP);
#pragma unroll
        for (int w = 0; w < WR; w++) {
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int p = 4 * w + b; // entry (j, k) = (p / 9, p % 9) of the row; 81..83: padding
                if (p < S * S) word |= ((run[p / S][(p % S) >> 2] >> (8 * ((p % S) & 3))) & 0xFFu) << (8 * b);
            }
            dst[w] = word;
        }
        if (i == 0) { // game padding (bytes 756..767) stays zero
            uint32_t *pad = reinterpret_cast<uint32_t *>(s_stage + q * GPITCH + S * RP);
            pad[0] = pad[1] = pad[2] = 0u;
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
        flags[gidx] = (uint8_t)((total == 0 ? TG_FLAG_TERMINAL : 0u) | (((badm >> (S * q)) & 0x1FFu) != 0 ? TG_FLAG_RANGE : 0u));
        nnz[gidx] = (int32_t)total;
        if (steps) steps[gidx] = my_steps;
    }
    if (lane == 0 && nwg > 0) {
        for (int qq = 0; qq < nwg; qq++) bulk_s2g(slab_out + (g0 + wg0 + qq) * GP, s_stage + qq * GPITCH, GP);
        bulk_commit();
        bulk_wait_read<0>();
    }
}

} // namespace roll9

int launch_rollout_rows9(const int8_t *slab_in, const uint8_t *tape, long long stride, int K, int8_t *slab_out, uint8_t *flags,
                         int32_t *nnz, int32_t *steps, long long B, int shift, int freeze, cudaStream_t st) {
    using namespace roll9;
    const long long grid = (B + TG - 1) / TG;
    if (grid > 0x7FFFFFFFLL) return TG_E_ARG;
    const int s3 = shift * shift * shift;
    const int chk = s3 >= 64 ? 1 : 64 / s3;
    if (freeze) {
        auto kern = rollout_rows9_kernel<true>;
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        kern<<<(int)grid, 32 * (NW + 1), SMEM_BYTES, st>>>(slab_in, tape, stride, K, slab_out, flags, nnz, steps, B, shift, chk);
    } else {
        auto kern = rollout_rows9_kernel<false>;
        TG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        kern<<<(int)grid, 32 * (NW + 1), SMEM_BYTES, st>>>(slab_in, tape, stride, K, slab_out, flags, nnz, steps, B, shift, chk);
    }
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

} // namespace tg
