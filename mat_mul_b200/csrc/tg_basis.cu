// tg_basis.cu -- K5: change-of-basis augmentation (C ABI: tg_change_of_basis,
// tg_change_of_basis_factors, tg_sample_unimodular).
//
// ABSENT from the reference (grep finds no basis/einsum code); specified from
// the AlphaTensor paper (Fawzi et al. 2022, Methods "Change of basis"):
//     T'[i][j][k] = sum_abc A[i][a] B[j][b] C[k][c] T[a][b][c]
//     u' = A u,  v' = B v,  w' = C w      (so sum u'(x)v'(x)w' = T')
// with A, B, C integer and unimodular.  Parity is pinned by our own int64
// oracle (oracle/tg_oracle.c orc_change_of_basis) and algebraic invariants.
//
// One CTA per game.  The three mode products are done as three passes of the
// SAME routine "contract the slowest axis, write the result rotated":
//     X[b][c][i] = sum_a A[i][a] T[a][b][c]
//     Y[c][i][j] = sum_b B[j][b] X[b][c][i]
//     Z[i][j][k] = sum_c C[k][c] Y[c][i][j]
// Intermediates live in shared memory as int32 (exact for any int8 input and
// int8 matrices); a thread owns one column of the contracted axis in registers
// and produces its S outputs.  TG_FLAG_RANGE marks games whose T' leaves the
// int8 slab's guaranteed zone [-64,63].
#include "tg_common.cuh"

namespace tg {

template <int S>
struct BasisCfg {
    using G = Geo<S>;
    static constexpr int S2 = S * S, S3 = S2 * S;
    static constexpr int NT = S2 >= 256 ? 256 : (S2 >= 64 ? 96 : 32);
    static constexpr int SMEM_BYTES = G::GP + 2 * S3 * 4 + 3 * S2 * 4 + 16;
};

// out[rest][a'] = sum_a M[a'][a] in[a][rest]; `rest` has S*S entries.
template <int S, int NT>
__device__ __forceinline__ void mode_pass(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                          const int32_t *__restrict__ M) {
    constexpr int S2 = S * S;
    for (int r = threadIdx.x; r < S2; r += NT) {
        int32_t col[S];
#pragma unroll
        for (int a = 0; a < S; a++) col[a] = in[a * S2 + r];
#pragma unroll
        for (int ap = 0; ap < S; ap++) {
            int32_t acc = 0;
#pragma unroll
            for (int a = 0; a < S; a++) acc += M[ap * S + a] * col[a];
            out[r * S + ap] = acc;
        }
    }
}

template <int S>
__global__ void __launch_bounds__(BasisCfg<S>::NT)
    basis_kernel(const int8_t *__restrict__ slab_in, const int8_t *__restrict__ mats, long long mat_stride,
                 int8_t *__restrict__ slab_out, uint8_t *__restrict__ flags, long long N) {
    using C = BasisCfg<S>;
    using G = Geo<S>;
    constexpr int NT = C::NT;
    extern __shared__ __align__(128) uint8_t smem[];
    int8_t *s_in = reinterpret_cast<int8_t *>(smem);
    int32_t *s_a = reinterpret_cast<int32_t *>(smem + G::GP);
    int32_t *s_b = s_a + C::S3;
    int32_t *s_m = s_b + C::S3; // A, B, C as int32 [3][S][S]
    uint32_t *s_flag = reinterpret_cast<uint32_t *>(s_m + 3 * C::S2);

    const int tid = threadIdx.x;
    const long long n = blockIdx.x;
    if (n >= N) return;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(slab_in + n * G::GP);
    for (int w = tid; w < G::GP / 4; w += NT) reinterpret_cast<uint32_t *>(s_in)[w] = src[w];
    const int8_t *m = mats + n * mat_stride;
    for (int q = tid; q < 3 * C::S2; q += NT) s_m[q] = (int32_t)m[q];
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    for (int e = tid; e < C::S3; e += NT) { // dense copy T[a][b][c]
        const int i = e / C::S2, jk = e % C::S2;
        s_a[e] = (int32_t)s_in[i * G::RP + jk];
    }
    __syncthreads();
    mode_pass<S, NT>(s_a, s_b, s_m); // contract a with A
    __syncthreads();
    mode_pass<S, NT>(s_b, s_a, s_m + C::S2); // contract b with B
    __syncthreads();
    mode_pass<S, NT>(s_a, s_b, s_m + 2 * C::S2); // contract c with C -> Z[i][j][k]
    __syncthreads();
    uint32_t bad = 0;
    uint32_t *dst = reinterpret_cast<uint32_t *>(slab_out + n * G::GP);
    for (int w = tid; w < G::GP / 4; w += NT) {
        const int i = w / G::WR, c = w % G::WR;
        uint32_t word = 0;
        if (i < S) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int jk = 4 * c + q;
                if (jk < C::S2) {
                    const int v = s_b[i * C::S2 + jk];
                    if (v < -64 || v > 63) bad = TG_FLAG_RANGE;
                    word |= ((uint32_t)v & 0xFFu) << (8 * q);
                }
            }
        }
        dst[w] = word;
    }
    if (bad) atomicOr(s_flag, bad);
    __syncthreads();
    if (tid == 0 && flags) flags[n] = (uint8_t)*s_flag;
}

// tokens of one game-step: coef' = M coef for the three factors; token' = coef' + shift_out
template <int S>
__global__ void basis_factors_kernel(const uint8_t *__restrict__ tape_in, long long in_step_stride, int shift_in,
                                     const int8_t *__restrict__ mats, long long mat_stride, uint8_t *__restrict__ tape_out,
                                     long long out_step_stride, int shift_out, uint8_t *__restrict__ flags, long long N,
                                     int R) {
    using G = Geo<S>;
    // thread = (game n, step r, output token q in [0, TP))
    const long long total = N * (long long)R * G::TP;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(idx % G::TP);
        const long long nr = idx / G::TP;
        const long long n = nr % N;
        const int r = (int)(nr / N);
        uint8_t tokv = 0;
        if (q < 3 * S) {
            const int f = q / S, i = q % S;
            const uint8_t *tin = tape_in + (size_t)r * in_step_stride + n * G::TP + f * S;
            const int8_t *M = mats + n * mat_stride + f * S * S + i * S;
            int acc = 0;
#pragma unroll
            for (int a = 0; a < S; a++) acc += (int)M[a] * ((int)tin[a] - shift_in);
            if (acc < -shift_out || acc > shift_out) flags[n] |= (uint8_t)TG_FLAG_TOKEN_RANGE; // same bit from every writer
            tokv = (uint8_t)(acc + shift_out);
        }
        tape_out[(size_t)r * out_step_stride + n * G::TP + q] = tokv;
    }
}

__device__ __forceinline__ void philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                             uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0, c1 = lo1, c2 = hi0 ^ c3 ^ k1, c3 = lo0;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

// one thread per (game, matrix): M = L * U, L unit-lower, U upper with diagonal +-1, off-diagonal entries
// -1/0/+1 with probabilities (p_nz/2, 1-p_nz, p_nz/2).  Draw for entry (r,c): byte (r*S+c)%16 of Philox block
// (r*S+c)/16 with ctr = (block, matrix f, 0x6D617473, d_lo), key as in demo generation.
template <int S>
__global__ void unimodular_kernel(unsigned long long seed, unsigned long long first, long long N, uint32_t thr_nz,
                                  int8_t *__restrict__ mats) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= N * 3) return;
    const long long n = t / 3;
    const int f = (int)(t % 3);
    const unsigned long long d = first + (unsigned long long)n;
    const uint32_t k0 = (uint32_t)seed ^ ((uint32_t)(d >> 32) * 0x9E3779B9u), k1 = (uint32_t)(seed >> 32);
    int8_t L[S * S], U[S * S];
    uint32_t blk[4] = {0, 0, 0, 0};
    for (int e = 0; e < S * S; e++) {
        if ((e & 15) == 0) philox_block((uint32_t)(e >> 4), (uint32_t)f, 0x6D617473u, (uint32_t)d, k0, k1, blk);
        const uint32_t byte = (blk[(e >> 2) & 3] >> (8 * (e & 3))) & 0xFFu;
        const int r = e / S, c = e % S;
        // low 7 bits: non-zero?  top bit: sign
        const int mag = (byte & 0x7Fu) < thr_nz ? 1 : 0;
        const int val = (byte & 0x80u) ? -mag : mag;
        L[e] = (int8_t)(r > c ? val : (r == c ? 1 : 0));
        U[e] = (int8_t)(r < c ? val : (r == c ? ((byte & 0x80u) ? -1 : 1) : 0));
    }
    int8_t *out = mats + (n * 3 + f) * S * S;
    for (int r = 0; r < S; r++)
        for (int c = 0; c < S; c++) {
            int acc = 0;
            for (int k = 0; k < S; k++) acc += (int)L[r * S + k] * (int)U[k * S + c];
            out[r * S + c] = (int8_t)acc;
        }
}

} // namespace tg

extern "C" {

int tg_change_of_basis(const int8_t *slab_in, const int8_t *mats, int per_game, int8_t *slab_out, uint8_t *flags, int64_t N,
                       int S, void *stream) {
    if (!tg::supported_S(S) || N < 0) return TG_E_ARG;
    if (N == 0) return TG_OK;
    if (!slab_in || !mats || !slab_out || slab_in == slab_out) return TG_E_ARG;
    if (((uintptr_t)slab_in | (uintptr_t)slab_out) & 15) return TG_E_ARG;
    if (N > 0x7FFFFFFFLL) return TG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long ms = per_game ? 3LL * S * S : 0;
    switch (S) {
    case 4: tg::basis_kernel<4><<<(int)N, tg::BasisCfg<4>::NT, tg::BasisCfg<4>::SMEM_BYTES, st>>>(slab_in, mats, ms, slab_out, flags, N); break;
    case 9: tg::basis_kernel<9><<<(int)N, tg::BasisCfg<9>::NT, tg::BasisCfg<9>::SMEM_BYTES, st>>>(slab_in, mats, ms, slab_out, flags, N); break;
    case 16: tg::basis_kernel<16><<<(int)N, tg::BasisCfg<16>::NT, tg::BasisCfg<16>::SMEM_BYTES, st>>>(slab_in, mats, ms, slab_out, flags, N); break; // 39 KB smem
    }
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int tg_change_of_basis_factors(const uint8_t *tape_in, int64_t in_step_stride, int shift_in, const int8_t *mats, int per_game,
                               uint8_t *tape_out, int64_t out_step_stride, int shift_out, uint8_t *flags, int64_t N, int R,
                               int S, void *stream) {
    if (!tg::supported_S(S) || N < 0 || R < 1 || shift_in < 0 || shift_out < 1 || shift_out > 127) return TG_E_ARG;
    if (N == 0) return TG_OK;
    if (!tape_in || !mats || !tape_out || !flags) return TG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long ms = per_game ? 3LL * S * S : 0;
    const long long total = N * (long long)R * ((3 * S + 15) & ~15);
    const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    switch (S) {
    case 4: tg::basis_factors_kernel<4><<<grid, 256, 0, st>>>(tape_in, in_step_stride, shift_in, mats, ms, tape_out, out_step_stride, shift_out, flags, N, R); break;
    case 9: tg::basis_factors_kernel<9><<<grid, 256, 0, st>>>(tape_in, in_step_stride, shift_in, mats, ms, tape_out, out_step_stride, shift_out, flags, N, R); break;
    case 16: tg::basis_factors_kernel<16><<<grid, 256, 0, st>>>(tape_in, in_step_stride, shift_in, mats, ms, tape_out, out_step_stride, shift_out, flags, N, R); break;
    }
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int tg_sample_unimodular(uint64_t seed, uint64_t first, int64_t N, int S, double p_nonzero, int8_t *mats, void *stream) {
    if (!tg::supported_S(S) || N < 0 || !(p_nonzero >= 0.0) || p_nonzero > 1.0) return TG_E_ARG;
    if (N == 0) return TG_OK;
    if (!mats) return TG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t thr = (uint32_t)(p_nonzero * 128.0);
    const int grid = (int)((N * 3 + 63) / 64);
    switch (S) {
    case 4: tg::unimodular_kernel<4><<<grid, 64, 0, st>>>(seed, first, N, thr, mats); break;
    case 9: tg::unimodular_kernel<9><<<grid, 64, 0, st>>>(seed, first, N, thr, mats); break;
    case 16: tg::unimodular_kernel<16><<<grid, 64, 0, st>>>(seed, first, N, thr, mats); break;
    }
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

} // extern "C"
