// tg_rows.cuh -- a ROW of a game as S runs of ceil(S / 4) packed words, shared by the row-owner kernels (tg_rollout9.cu,
// tg_expand9.cu).
//
// Row i of a game is the S*S entries (i, j, k) = RP bytes of the slab.  A row-owner thread keeps it as run[j][m]: run j holds
// the entries (i, j, 0..S-1) as packed bytes, 4 + 4 + 1 for 9x9x9 (bytes 1..3 of word 2 are padding), four aligned words for
// 16x16x16.  In that form the rank-1 update needs no pack(v w) word that straddles two j:
//     run[j][m] += (-u_i v_j) * W_m ,   W_m = the words of pack(w) in integer form.
// unpack_row / pack_row convert between the slab's row words (runs of S bytes back to back) and the runs; both are straight-
// line code (funnel shifts in, at most two PRMTs per word out) with compile-time selectors.
#pragma once
#include "tg_common.cuh"

namespace tg {

template <int S>
struct RowRuns {
    static constexpr int KW = (S + 3) / 4;         // words per run
    static constexpr int WR = Geo<S>::RP / 4;      // words per slab row
    static constexpr int S2 = S * S;
    // source of row byte p (< S*S): run word index 4 * j + m flattened as j * KW + m, and the byte inside it
    static constexpr __host__ __device__ int src_word(int p) { return (p / S) * KW + ((p % S) >> 2); }
    static constexpr __host__ __device__ int src_byte(int p) { return (p % S) & 3; }
    // the (at most three) distinct run words that feed slab word w, in order of appearance; -1 = none
    static constexpr __host__ __device__ int feed(int w, int which) {
        int found[3] = {-1, -1, -1}, n = 0;
        for (int b = 0; b < 4; b++) {
            const int p = 4 * w + b;
            if (p >= S2) break;
            const int sw = src_word(p);
            bool seen = false;
            for (int x = 0; x < n; x++) seen = seen || found[x] == sw;
            if (!seen && n < 3) found[n++] = sw;
        }
        return found[which];
    }
    // selector of the first PRMT (sources feed 0, feed 1) and of the second one (its result, feed 2)
    static constexpr __host__ __device__ uint32_t sel1(int w) {
        uint32_t sel = 0;
        for (int b = 0; b < 4; b++) {
            const int p = 4 * w + b;
            uint32_t nib = 0;
            if (p < S2) {
                const int sw = src_word(p);
                if (sw == feed(w, 0)) nib = (uint32_t)src_byte(p);
                else if (sw == feed(w, 1)) nib = 4u + (uint32_t)src_byte(p);
            }
            sel |= nib << (4 * b);
        }
        return sel;
    }
    static constexpr __host__ __device__ uint32_t sel2(int w) {
        uint32_t sel = 0;
        for (int b = 0; b < 4; b++) {
            const int p = 4 * w + b;
            uint32_t nib = (uint32_t)b; // keep the byte of the first PRMT
            if (p < S2 && src_word(p) == feed(w, 2)) nib = 4u + (uint32_t)src_byte(p);
            sel |= nib << (4 * b);
        }
        return sel;
    }
    static constexpr __host__ __device__ uint32_t valid_mask(int w) { // bytes of slab word w that are entries of the row
        uint32_t m = 0;
        for (int b = 0; b < 4; b++)
            if (4 * w + b < S2) m |= 0xFFu << (8 * b);
        return m;
    }
};

__device__ __forceinline__ uint32_t prmt_rows(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// slab row words r[0 .. WR-1] (two's complement bytes; r must have two spare elements) -> runs (two's complement bytes, the
// unused bytes of a run's last word zero)
template <int S>
__device__ __forceinline__ void unpack_row(uint32_t (&r)[RowRuns<S>::WR + 2], uint32_t (&run)[S][RowRuns<S>::KW]) {
    using R = RowRuns<S>;
    if constexpr (R::S2 % 4 != 0) r[R::WR - 1] &= 0xFFFFFFFFu >> (8 * (4 - R::S2 % 4)); // row padding
    r[R::WR] = r[R::WR + 1] = 0;
#pragma unroll
    for (int j = 0; j < S; j++) {
        const int o = S * j, w0 = o >> 2, sh = 8 * (o & 3); // the run starts at byte S j of the row
#pragma unroll
        for (int m = 0; m < R::KW; m++) {
            uint32_t x = __funnelshift_r(r[w0 + m], r[w0 + m + 1], sh);
            if (S - 4 * m < 4) x &= 0xFFFFFFFFu >> (8 * (4 - (S - 4 * m)));
            run[j][m] = x;
        }
    }
}

// runs (two's complement bytes; the unused bytes of a run's last word may hold anything) -> slab row words, padding zero.
// One word per template instance so that the feeds and selectors are constant expressions (a loop variable, even of an
// unrolled loop, would leave the run indices to the optimiser -- and the runs in local memory).
template <int S, int W>
__device__ __forceinline__ void pack_word(const uint32_t (&run)[S][RowRuns<S>::KW], uint32_t (&out)[RowRuns<S>::WR]) {
    using R = RowRuns<S>;
    constexpr int a = R::feed(W, 0), b = R::feed(W, 1), c = R::feed(W, 2);
    constexpr uint32_t s1 = R::sel1(W), s2 = R::sel2(W), vm = R::valid_mask(W);
    if constexpr (a < 0) {
        out[W] = 0;
    } else {
        uint32_t x;
        if constexpr (b >= 0) x = prmt_rows(run[a / R::KW][a % R::KW], run[b / R::KW][b % R::KW], s1);
        else x = prmt_rows(run[a / R::KW][a % R::KW], 0u, s1);
        if constexpr (c >= 0) x = prmt_rows(x, run[c / R::KW][c % R::KW], s2);
        out[W] = vm == 0xFFFFFFFFu ? x : (x & vm);
    }
    if constexpr (W + 1 < R::WR) pack_word<S, W + 1>(run, out);
}
template <int S>
__device__ __forceinline__ void pack_row(const uint32_t (&run)[S][RowRuns<S>::KW], uint32_t (&out)[RowRuns<S>::WR]) {
    pack_word<S, 0>(run, out);
}

} // namespace tg
