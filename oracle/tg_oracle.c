/*
 * tg_oracle.c -- CPU ORACLE for the TensorGame hot path of kurtosis/mat_mul.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT.  It is a plain-C restatement
 * of the reference's per-game Python/PyTorch arithmetic, written from the
 * behaviour of the reference (file:line cited on every function) and pinned
 * against the live reference through tests/golden/ (see oracle/gen_golden.py).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (mat_mul_b200/) never does.
 *
 * Conventions: residual tensors are dense int32 [S][S][S], index (i*S+j)*S+k
 * with u -> i, v -> j, w -> k (utils.py:69-85).  Action tokens are int32
 * [3S] = cat(u, v, w) + shift (utils.py:56-66, 231).  All arithmetic on the
 * path is exact integer arithmetic; the reference carries it in float32.
 *
 * Third-party pieces the reference leans on (not vendored in /root/reference):
 *   - PyTorch 2.11.0 CPU generator: MT19937 + 53-bit doubles
 *     (at::CPUGeneratorImpl::random64, at::uniform_real_distribution<double>)
 *   - PyTorch 2.11.0 torch.multinomial CPU kernel (with replacement) and
 *     torch.distributions.Categorical normalisation
 *   - torch.linalg.matrix_rank (LAPACK gesdd, float32) -- restated as an exact
 *     rank over the rationals; see orc_slice_rank.
 * Their published algorithms are restated below and pinned by golden vectors
 * generated from the live torch in gen_golden.py.
 *
 * PARITY PINNING: steps, demos, Strassen, matmul tensor, getitem, ranks are
 * pinned to reference outputs (tests/golden/).  Change of basis and the Philox
 * demo stream have NO reference counterpart: "parity unpinned" for those two
 * (pinned only by Random123 known-answer vectors and algebraic invariants).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_FLAG_TERMINAL 1u /* new head is all zero: utils.py:181-188 on the head, act.py:177 */
#define ORC_FLAG_NULL 2u     /* rank-1 update was all zero: utils.py:191-194 */

/* ------------------------------------------------------------------ */
/* RNG: torch CPU generator = MT19937 (at::mt19937), init_genrand seed */
/* ------------------------------------------------------------------ */
typedef struct {
    uint32_t s[624];
    int pos;
} orc_mt;

void orc_mt_seed(orc_mt *g, uint32_t seed) {
    g->s[0] = seed;
    for (int i = 1; i < 624; i++)
        g->s[i] = 1812433253u * (g->s[i - 1] ^ (g->s[i - 1] >> 30)) + (uint32_t)i;
    g->pos = 624;
}

static void orc_mt_twist(orc_mt *g) {
    uint32_t *s = g->s;
    for (int i = 0; i < 624; i++) {
        uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
        uint32_t x = s[(i + 397) % 624] ^ (y >> 1);
        if (y & 1u) x ^= 0x9908b0dfu;
        s[i] = x;
    }
    g->pos = 0;
}

uint32_t orc_mt_u32(orc_mt *g) {
    if (g->pos >= 624) orc_mt_twist(g);
    uint32_t y = g->s[g->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* torch.rand(dtype=float64) on CPU: random64() = (first draw << 32) | second
 * draw, keep the low 53 bits, scale by 2^-53.  Probed against torch 2.11. */
double orc_mt_f64(orc_mt *g) {
    uint64_t hi = orc_mt_u32(g);
    uint64_t lo = orc_mt_u32(g);
    uint64_t r = ((hi << 32) | lo) & ((1ull << 53) - 1);
    return (double)r * (1.0 / 9007199254740992.0);
}

void orc_mt_fill_f64(uint32_t seed, int64_t n, double *out) {
    orc_mt g;
    orc_mt_seed(&g, seed);
    for (int64_t i = 0; i < n; i++) out[i] = orc_mt_f64(&g);
}

/* ------------------------------------------------------------------ */
/* Categorical(probs).sample == torch.multinomial(probs/sum, n, True)  */
/* utils.py:197-200, datasets.py:155-158                               */
/* ------------------------------------------------------------------ */
/* cdf[] as the multinomial CPU kernel builds it: float32 running sum of the
 * Categorical-normalised probabilities, divided by the total, last := 1. */
void orc_categorical_cdf(const float *probs, int n, float *cdf) {
    float total = 0.f;
    for (int i = 0; i < n; i++) total += probs[i];
    float run = 0.f;
    for (int i = 0; i < n; i++) {
        float p = probs[i] / total; /* Categorical.__init__: probs / probs.sum(-1) */
        run += p;
        cdf[i] = run;
    }
    float last = run;
    for (int i = 0; i < n; i++) cdf[i] /= last;
    cdf[n - 1] = 1.f;
}

/* first bucket whose cdf >= u (binary search of the kernel, double compare) */
int orc_categorical_pick(const float *cdf, int n, double u) {
    int lo = 0, hi = n;
    while (hi - lo > 0) {
        int mid = lo + (hi - lo) / 2;
        if ((double)cdf[mid] < u)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo < n ? lo : n - 1;
}

/* ------------------------------------------------------------------ */
/* env primitives                                                      */
/* ------------------------------------------------------------------ */
/* utils.py:56-66 */
void orc_action_to_uvw(const int32_t *action, int S, int shift, int32_t *u, int32_t *v, int32_t *w) {
    for (int i = 0; i < S; i++) {
        u[i] = action[i] - shift;
        v[i] = action[S + i] - shift;
        w[i] = action[2 * S + i] - shift;
    }
}

/* utils.py:69-85 : out[i][j][k] = u[i] v[j] w[k] */
void orc_uvw_to_tensor(const int32_t *u, const int32_t *v, const int32_t *w, int S, int32_t *out) {
    for (int i = 0; i < S; i++)
        for (int j = 0; j < S; j++)
            for (int k = 0; k < S; k++) out[(i * S + j) * S + k] = u[i] * v[j] * w[k];
}

/* utils.py:88-96 when shift == 1 (the reference hard-codes 1);
 * explicit shift for every other caller */
void orc_action_to_tensor(const int32_t *action, int S, int shift, int32_t *out) {
    int32_t u[64], v[64], w[64];
    orc_action_to_uvw(action, S, shift, u, v, w);
    orc_uvw_to_tensor(u, v, w, S, out);
}

/* One transition of one game.
 *   new_head = head - u(x)v(x)w          act.py:266-275, training.py:253-255
 *   TERMINAL = (new_head == 0).all()     utils.py:181-188 applied to the head (act.py:177)
 *   NULL     = !(new_head != head).any() utils.py:191-194
 *   nnz      = count(new_head != 0)      training.py:259-266 (rank upper bound)
 */
void orc_step(const int32_t *T_in, const int32_t *action, int S, int shift, int32_t *T_out,
              uint32_t *flags, int32_t *nnz) {
    int32_t u[64], v[64], w[64];
    orc_action_to_uvw(action, S, shift, u, v, w);
    int n = 0, changed = 0;
    for (int i = 0; i < S; i++)
        for (int j = 0; j < S; j++)
            for (int k = 0; k < S; k++) {
                int idx = (i * S + j) * S + k;
                int32_t d = u[i] * v[j] * w[k];
                int32_t t = T_in[idx] - d;
                changed |= (d != 0);
                n += (t != 0);
                T_out[idx] = t;
            }
    *flags = (n == 0 ? ORC_FLAG_TERMINAL : 0u) | (changed ? 0u : ORC_FLAG_NULL);
    *nnz = n;
}

void orc_step_batch(const int32_t *T_in, const int32_t *actions, int64_t B, int S, int shift,
                    int32_t *T_out, uint32_t *flags, int32_t *nnz) {
    const int64_t S3 = (int64_t)S * S * S;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++)
        orc_step(T_in + b * S3, actions + b * 3 * S, S, shift, T_out + b * S3, flags + b, nnz + b);
}

/* The reference's own data types (float32 residuals, int64 tokens) for the
 * CPU baseline leg of bench.py: training.py:253-266 per game, threaded over
 * games.  Same arithmetic as orc_step. */
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
void orc_step_batch_f32(const float *T_in, const int64_t *actions, int64_t B, int S, int shift,
                        float *T_out, uint8_t *flags, int32_t *nnz) {
    const int64_t S3 = (int64_t)S * S * S;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) {
        const float *tin = T_in + b * S3;
        float *tout = T_out + b * S3;
        const int64_t *a = actions + b * 3 * S;
        float wf[64];
        int unz = 0, vnz = 0, wnz = 0;
        for (int k = 0; k < S; k++) {
            wf[k] = (float)(a[2 * S + k] - shift);
            unz |= (a[k] != shift), vnz |= (a[S + k] != shift), wnz |= (a[2 * S + k] != shift);
        }
        int n = 0;
        for (int i = 0; i < S; i++) {
            const float ui = (float)(a[i] - shift);
            for (int j = 0; j < S; j++) {
                const float uv = ui * (float)(a[S + j] - shift);
                const float *ti = tin + (i * S + j) * S;
                float *to = tout + (i * S + j) * S;
                int nn = 0;
                for (int k = 0; k < S; k++) {
                    const float t = ti[k] - uv * wf[k];
                    nn += (t != 0.f);
                    to[k] = t;
                }
                n += nn;
            }
        }
        flags[b] = (uint8_t)((n == 0 ? ORC_FLAG_TERMINAL : 0u) | ((unz && vnz && wnz) ? 0u : ORC_FLAG_NULL));
        nnz[b] = n;
    }
}

/* datasets.py:144-153 : T <- T - action_to_tensor(a) for a in action list */
void orc_take_actions(const int32_t *actions, int n_actions, int S, int shift, int32_t *T) {
    int32_t tmp[4096];
    const int S3 = S * S * S;
    for (int a = 0; a < n_actions; a++) {
        orc_action_to_tensor(actions + a * 3 * S, S, shift, tmp);
        for (int e = 0; e < S3; e++) T[e] -= tmp[e];
    }
}

/* K-step replay of one game with per-step bookkeeping (fused-rollout oracle).
 * tape is [K][3S].  After the head first becomes all zero the game is frozen
 * (later tape entries are ignored), mirroring the break at act.py:49.
 * steps_out = number of actions applied; ret_out = -steps (one -1 per action,
 * act.py:60-62) -- the terminal -get_rank term is added by the caller. */
void orc_rollout(const int32_t *T_in, const int32_t *tape, int K, int S, int shift, int32_t *T_out,
                 uint32_t *flags, int32_t *nnz, int32_t *steps_out) {
    const int S3 = S * S * S;
    int32_t cur[4096], nxt[4096];
    memcpy(cur, T_in, sizeof(int32_t) * S3);
    int n = 0;
    for (int e = 0; e < S3; e++) n += (cur[e] != 0);
    uint32_t f = n == 0 ? ORC_FLAG_TERMINAL : 0u;
    int steps = 0;
    for (int t = 0; t < K && !(f & ORC_FLAG_TERMINAL); t++) {
        uint32_t sf;
        orc_step(cur, tape + t * 3 * S, S, shift, nxt, &sf, &n);
        memcpy(cur, nxt, sizeof(int32_t) * S3);
        f = sf & ORC_FLAG_TERMINAL;
        steps++;
    }
    memcpy(T_out, cur, sizeof(int32_t) * S3);
    *flags = f;
    *nnz = n;
    *steps_out = steps;
}

void orc_rollout_batch(const int32_t *T_in, const int32_t *tape, int64_t B, int K, int S, int shift,
                       int32_t *T_out, uint32_t *flags, int32_t *nnz, int32_t *steps_out) {
    const int64_t S3 = (int64_t)S * S * S;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++)
        orc_rollout(T_in + b * S3, tape + b * K * 3 * S, K, S, shift, T_out + b * S3, flags + b,
                    nnz + b, steps_out + b);
}

/* ------------------------------------------------------------------ */
/* synthetic demonstrations                                            */
/* utils.py:203-233 == datasets.py:124-142                             */
/* ------------------------------------------------------------------ */
/* Consume the uniform stream exactly as the reference's loop does: each try
 * draws S doubles for u, S for v, S for w (three Categorical.sample calls),
 * the try is rejected iff u(x)v(x)w is all zero, i.e. iff one of u, v, w is
 * all zero; accepted tries are appended in order, R per demo, demos back to
 * back.  Returns the number of COMPLETE demos produced; *consumed = doubles
 * used by those demos.  tokens_out [n][R][3S], targets_out [n][S^3]. */
int64_t orc_demos_from_ustream(const double *ustream, int64_t n_u, const int32_t *values,
                               const float *probs, int n_values, int R, int S, int shift,
                               int64_t n_demos, int32_t *tokens_out, int32_t *targets_out,
                               int64_t *consumed) {
    float cdf[64];
    orc_categorical_cdf(probs, n_values, cdf);
    const int S3 = S * S * S;
    int64_t pos = 0, done = 0, used = 0;
    int32_t tmp[4096];
    for (int64_t d = 0; d < n_demos; d++) {
        int32_t *tgt = targets_out + d * S3;
        memset(tgt, 0, sizeof(int32_t) * S3);
        int ok = 1;
        for (int r = 0; r < R && ok; r++) {
            for (;;) {
                if (pos + 3 * S > n_u) {
                    ok = 0;
                    break;
                }
                int32_t f[3][64];
                int allzero[3] = {1, 1, 1};
                for (int m = 0; m < 3; m++)
                    for (int i = 0; i < S; i++) {
                        int32_t val = values[orc_categorical_pick(cdf, n_values, ustream[pos++])];
                        f[m][i] = val;
                        if (val != 0) allzero[m] = 0;
                    }
                if (allzero[0] || allzero[1] || allzero[2]) continue; /* utils.py:229 */
                int32_t *tok = tokens_out + (d * R + r) * 3 * S;
                for (int m = 0; m < 3; m++)
                    for (int i = 0; i < S; i++) tok[m * S + i] = f[m][i] + shift;
                orc_uvw_to_tensor(f[0], f[1], f[2], S, tmp);
                for (int e = 0; e < S3; e++) tgt[e] += tmp[e];
                break;
            }
        }
        if (!ok) break;
        done++;
        used = pos;
    }
    if (consumed) *consumed = used;
    return done;
}

/* ------------------------------------------------------------------ */
/* SyntheticDemoDataset.__getitem__  datasets.py:77-122                */
/* ------------------------------------------------------------------ */
/* tokens [R][3S], target [S^3] -> state [dim_t][S^3], scalar, action, reward.
 * replay_shift is the shift used by action_to_tensor inside the reference,
 * which is always 1; callers wanting the corrected behaviour pass
 * the dataset's shift. */
void orc_demo_getitem(const int32_t *tokens, const int32_t *target, int R, int S, int dim_t,
                      int idx_action, int replay_shift, int32_t *state_out, float *scalar_out,
                      int32_t *action_out, float *reward_out) {
    const int S3 = S * S * S;
    memset(state_out, 0, sizeof(int32_t) * S3 * dim_t);
    memcpy(state_out, target, sizeof(int32_t) * S3);
    if (idx_action != R - 1) /* datasets.py:90-92 */
        orc_take_actions(tokens + (idx_action + 1) * 3 * S, R - 1 - idx_action, S, replay_shift,
                         state_out);
    /* datasets.py:94-104 : reversed(action_seq[idx+1 : idx+dim_t]) */
    int lo = idx_action + 1, hi = idx_action + dim_t;
    if (hi > R) hi = R;
    int slot = 1;
    for (int a = hi - 1; a >= lo; a--, slot++)
        orc_action_to_tensor(tokens + a * 3 * S, S, replay_shift, state_out + slot * S3);
    *scalar_out = (float)(R - idx_action);    /* datasets.py:115 */
    *reward_out = -(float)(idx_action + 1);   /* datasets.py:116 */
    memcpy(action_out, tokens + idx_action * 3 * S, sizeof(int32_t) * 3 * S);
}

/* ------------------------------------------------------------------ */
/* Strassen demo + matmul tensor                                       */
/* ------------------------------------------------------------------ */
/* Strassen's seven products M1..M7 for C = A B with 2x2 row-major vec():
 * rows are the u (A side), v (B side), w (C side) factors in the order the
 * reference lists them, datasets.py:423-460.  Written as +/0/- strings. */
static const char *const STRASSEN_U[7] = {"+00+", "00++", "+000", "000+", "++00", "-0+0", "0+0-"};
static const char *const STRASSEN_V[7] = {"+00+", "+000", "0+0-", "-0+0", "000+", "++00", "00++"};
static const char *const STRASSEN_W[7] = {"+00+", "00+-", "0+0+", "+0+0", "-+00", "000+", "+000"};

static int32_t sgn(char c) { return c == '+' ? 1 : (c == '-' ? -1 : 0); }

/* uu, vv, ww: [7][4] */
void orc_strassen_factors(int32_t *uu, int32_t *vv, int32_t *ww) {
    for (int r = 0; r < 7; r++)
        for (int i = 0; i < 4; i++) {
            uu[r * 4 + i] = sgn(STRASSEN_U[r][i]);
            vv[r * 4 + i] = sgn(STRASSEN_V[r][i]);
            ww[r * 4 + i] = sgn(STRASSEN_W[r][i]);
        }
}

/* utils.py:40-53 : sum of n rank-1 terms into a (S,S,S) tensor and the token
 * list cat(u,v,w)+shift.  (The reference hard-codes S=4) */
void orc_uvw_to_demo(const int32_t *uu, const int32_t *vv, const int32_t *ww, int n, int S, int shift,
                     int32_t *tensor_out, int32_t *actions_out) {
    const int S3 = S * S * S;
    int32_t tmp[4096];
    memset(tensor_out, 0, sizeof(int32_t) * S3);
    for (int r = 0; r < n; r++) {
        orc_uvw_to_tensor(uu + r * S, vv + r * S, ww + r * S, S, tmp);
        for (int e = 0; e < S3; e++) tensor_out[e] += tmp[e];
        for (int i = 0; i < S; i++) {
            actions_out[r * 3 * S + i] = uu[r * S + i] + shift;
            actions_out[r * 3 * S + S + i] = vv[r * S + i] + shift;
            actions_out[r * 3 * S + 2 * S + i] = ww[r * S + i] + shift;
        }
    }
}

/* StrassenDemoDataset.__init__  datasets.py:370-408.  Emits, in the
 * reference's order, for every 7-bit string (MSB = factor 0) one item per
 * unused factor: state = strassen - sum(used), action = factor + 2,
 * reward = -n_avail, scalar = 0.  Returns the item count (448).
 * states [448][64], actions [448][12], rewards [448], bits [448]. */
int orc_strassen_dataset(int32_t *states, int32_t *actions, float *rewards, int32_t *bits) {
    int32_t uu[28], vv[28], ww[28], full[64], toks[84], tmp[64];
    orc_strassen_factors(uu, vv, ww);
    orc_uvw_to_demo(uu, vv, ww, 7, 4, 1, full, toks);
    int n = 0;
    for (int code = 0; code < 128; code++) {
        int32_t cur[64];
        memcpy(cur, full, sizeof(cur));
        int n_avail = 0;
        for (int r = 0; r < 7; r++) {
            int used = (code >> (6 - r)) & 1; /* format(i,"b").zfill(7)[r] */
            if (used) {
                orc_uvw_to_tensor(uu + r * 4, vv + r * 4, ww + r * 4, 4, tmp);
                for (int e = 0; e < 64; e++) cur[e] -= tmp[e];
            } else
                n_avail++;
        }
        for (int r = 0; r < 7; r++) {
            if ((code >> (6 - r)) & 1) continue;
            memcpy(states + n * 64, cur, sizeof(cur));
            for (int i = 0; i < 4; i++) {
                actions[n * 12 + i] = uu[r * 4 + i] + 2; /* datasets.py:397 */
                actions[n * 12 + 4 + i] = vv[r * 4 + i] + 2;
                actions[n * 12 + 8 + i] = ww[r * 4 + i] + 2;
            }
            rewards[n] = -(float)n_avail;
            bits[n] = code;
            n++;
        }
    }
    return n;
}

/* utils.py:143-161 for square n x n matrices: slot-0 tensor of shape
 * (n^2, n^2, n^2) with T[(ik/n)*n + j][j*n + ik%n][ik] = 1. */
void orc_build_matmul_tensor(int n, int32_t *out) {
    const int S = n * n;
    memset(out, 0, sizeof(int32_t) * S * S * S);
    for (int ik = 0; ik < n * n; ik++)
        for (int j = 0; j < n; j++) {
            int a = (ik / n) * n + j, b = j * n + ik % n;
            out[(a * S + b) * S + ik] = 1;
        }
}

/* ------------------------------------------------------------------ */
/* get_rank  utils.py:134-140 : sum over slices T[i,:,:] of matrix rank */
/* ------------------------------------------------------------------ */
/* The reference uses a float32 SVD with the default tolerance.  For the small
 * integer matrices of the game the numerical rank equals the exact rank over
 * the rationals; the oracle computes the exact rank by Gaussian elimination
 * modulo p = 2^61-1 with 128-bit products, which can only under-count when p
 * divides a pivot minor (probability ~ S/p).  Pinned against
 * torch.linalg.matrix_rank through tests/golden/ranks.npz. */
static uint64_t mulmod61(uint64_t a, uint64_t b) {
    const uint64_t P = (1ull << 61) - 1;
    __uint128_t z = (__uint128_t)a * b;
    uint64_t lo = (uint64_t)(z & P), hi = (uint64_t)(z >> 61);
    uint64_t r = lo + hi;
    if (r >= P) r -= P;
    return r;
}
static uint64_t powmod61(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = mulmod61(r, a);
        a = mulmod61(a, a);
        e >>= 1;
    }
    return r;
}

int orc_matrix_rank(const int32_t *M, int rows, int cols) {
    const uint64_t P = (1ull << 61) - 1;
    uint64_t a[64 * 64];
    for (int i = 0; i < rows * cols; i++) {
        int64_t v = M[i];
        a[i] = v >= 0 ? (uint64_t)v % P : P - ((uint64_t)(-v) % P);
        if (a[i] == P) a[i] = 0;
    }
    int rank = 0;
    for (int c = 0; c < cols && rank < rows; c++) {
        int piv = -1;
        for (int r = rank; r < rows; r++)
            if (a[r * cols + c]) {
                piv = r;
                break;
            }
        if (piv < 0) continue;
        if (piv != rank)
            for (int k = 0; k < cols; k++) {
                uint64_t t = a[piv * cols + k];
                a[piv * cols + k] = a[rank * cols + k];
                a[rank * cols + k] = t;
            }
        uint64_t inv = powmod61(a[rank * cols + c], P - 2);
        for (int r = rank + 1; r < rows; r++) {
            uint64_t f = mulmod61(a[r * cols + c], inv);
            if (!f) continue;
            for (int k = c; k < cols; k++) {
                uint64_t s = mulmod61(f, a[rank * cols + k]);
                uint64_t t = a[r * cols + k] + P - s;
                a[r * cols + k] = t >= P ? t - P : t;
            }
        }
        rank++;
    }
    return rank;
}

int orc_slice_rank(const int32_t *T, int S) {
    int total = 0;
    for (int i = 0; i < S; i++) total += orc_matrix_rank(T + i * S * S, S, S);
    return total;
}

void orc_slice_rank_batch(const int32_t *T, int64_t B, int S, int32_t *ranks) {
    const int64_t S3 = (int64_t)S * S * S;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) ranks[b] = orc_slice_rank(T + b * S3, S);
}

/* ------------------------------------------------------------------ */
/* change of basis  T' = T x1 A x2 B x3 C  (ABSENT from the reference; */
/* AlphaTensor paper, Methods "Change of basis") -- PARITY UNPINNED    */
/* ------------------------------------------------------------------ */
/* T'[i][j][k] = sum_abc A[i][a] B[j][b] C[k][c] T[a][b][c];  int64 inside. */
void orc_change_of_basis(const int32_t *T, const int32_t *A, const int32_t *B, const int32_t *C, int S,
                         int64_t *out) {
    const int S3 = S * S * S;
    int64_t *t1 = (int64_t *)malloc(sizeof(int64_t) * S3);
    int64_t *t2 = (int64_t *)malloc(sizeof(int64_t) * S3);
    for (int i = 0; i < S; i++)
        for (int b = 0; b < S; b++)
            for (int c = 0; c < S; c++) {
                int64_t acc = 0;
                for (int a = 0; a < S; a++) acc += (int64_t)A[i * S + a] * T[(a * S + b) * S + c];
                t1[(i * S + b) * S + c] = acc;
            }
    for (int i = 0; i < S; i++)
        for (int j = 0; j < S; j++)
            for (int c = 0; c < S; c++) {
                int64_t acc = 0;
                for (int b = 0; b < S; b++) acc += (int64_t)B[j * S + b] * t1[(i * S + b) * S + c];
                t2[(i * S + j) * S + c] = acc;
            }
    for (int i = 0; i < S; i++)
        for (int j = 0; j < S; j++)
            for (int k = 0; k < S; k++) {
                int64_t acc = 0;
                for (int c = 0; c < S; c++) acc += (int64_t)C[k * S + c] * t2[(i * S + j) * S + c];
                out[(i * S + j) * S + k] = acc;
            }
    free(t1);
    free(t2);
}

/* factors map u -> A u, v -> B v, w -> C w ; factors [R][3S] (values, not tokens) */
void orc_change_of_basis_factors(const int32_t *factors, int R, const int32_t *A, const int32_t *B,
                                 const int32_t *C, int S, int64_t *out) {
    const int32_t *M[3] = {A, B, C};
    for (int r = 0; r < R; r++)
        for (int m = 0; m < 3; m++)
            for (int i = 0; i < S; i++) {
                int64_t acc = 0;
                for (int a = 0; a < S; a++) acc += (int64_t)M[m][i * S + a] * factors[(r * 3 + m) * S + a];
                out[(r * 3 + m) * S + i] = acc;
            }
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as    */
/* easy as 1, 2, 3"; Random123 v1.14 kat_vectors) -- device RNG        */
/* contract for throughput-mode demo generation.  PARITY UNPINNED by   */
/* the reference (it has no device RNG); pinned by Random123 vectors.  */
/* ------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

/* Integer CDF for the device sampler: thr[i] = floor(cdf_i * 2^15) with the
 * float64 running sum of probs/sum; a 15-bit draw x picks the first i < n-1
 * with x < thr[i]; the last bucket catches everything (32768 = never). */
void orc_philox_thresholds(const double *probs, int n, uint32_t *thr) {
    double total = 0, run = 0;
    for (int i = 0; i < n; i++) total += probs[i];
    for (int i = 0; i < n; i++) {
        run += probs[i] / total;
        double t = run * 32768.0;
        thr[i] = (i == n - 1 || t >= 32768.0) ? 32768u : (uint32_t)t;
    }
}

/* Throughput-mode demo contract (ours).  For demo d (global index), term r,
 * try t, the 3S coefficient draws q = 0..3S-1 are 15-bit values of the Philox
 * stream: block bq = q / 8 with
 *   ctr = (d_lo , d_hi , r | t << 16 , bq) , key = (seed_lo , seed_hi),
 * word (q / 2) % 4 of the block, bits 0-14 for even q, bits 16-30 for odd q;
 * draw x -> first i < n-1 with x < thr[i], else n-1.  A try is rejected iff u,
 * v or w is all zero (utils.py:229); at most max_tries tries, after which the
 * term is forced to the unit triple (u=v=w=e_0 * values[n-1]) and *exhausted
 * is incremented (the reference would spin forever).  Results are
 * independent of how demos are partitioned over ranks. */
void orc_demo_philox(uint64_t seed, uint64_t d, const int32_t *values, const uint32_t *thr, int n_values,
                     int R, int S, int shift, int max_tries, int32_t *tokens_out, int32_t *target_out,
                     int32_t *exhausted) {
    const int S3 = S * S * S;
    int32_t tmp[4096];
    memset(target_out, 0, sizeof(int32_t) * S3);
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < R; r++) {
        int32_t f[3][64];
        int ok = 0;
        for (int t = 0; t < max_tries && !ok; t++) {
            int allzero[3] = {1, 1, 1};
            uint32_t blk[4];
            for (int q = 0; q < 3 * S; q++) {
                if ((q & 7) == 0) {
                    uint32_t ctr[4] = {(uint32_t)d, (uint32_t)(d >> 32), (uint32_t)r | ((uint32_t)t << 16), (uint32_t)(q >> 3)};
                    orc_philox4x32_10(ctr, key, blk);
                }
                uint32_t word = blk[(q >> 1) & 3];
                uint32_t x = ((q & 1) ? (word >> 16) : word) & 0x7FFFu;
                int i = 0;
                while (i < n_values - 1 && x >= thr[i]) i++;
                int32_t val = values[i];
                f[q / S][q % S] = val;
                if (val != 0) allzero[q / S] = 0;
            }
            ok = !(allzero[0] || allzero[1] || allzero[2]);
        }
        if (!ok) {
            for (int m = 0; m < 3; m++)
                for (int i = 0; i < S; i++) f[m][i] = (i == 0) ? values[n_values - 1] : 0;
            if (exhausted) (*exhausted)++;
        }
        int32_t *tok = tokens_out + r * 3 * S;
        for (int m = 0; m < 3; m++)
            for (int i = 0; i < S; i++) tok[m * S + i] = f[m][i] + shift;
        orc_uvw_to_tensor(f[0], f[1], f[2], S, tmp);
        for (int e = 0; e < S3; e++) target_out[e] += tmp[e];
    }
}

void orc_demos_philox_batch(uint64_t seed, uint64_t d0, int64_t n, const int32_t *values,
                            const uint32_t *thr, int n_values, int R, int S, int shift, int max_tries,
                            int32_t *tokens_out, int32_t *targets_out, int32_t *exhausted) {
    const int64_t S3 = (int64_t)S * S * S;
    int32_t ex = 0;
#pragma omp parallel for schedule(static) reduction(+ : ex)
    for (int64_t i = 0; i < n; i++) {
        int32_t e = 0;
        orc_demo_philox(seed, d0 + (uint64_t)i, values, thr, n_values, R, S, shift, max_tries,
                        tokens_out + i * R * 3 * S, targets_out + i * S3, &e);
        ex += e;
    }
    if (exhausted) *exhausted = ex;
}

/* ---- Throughput-mode demo contract v2 (ours): the "group alias" sampler.
 *
 * The reference draws every entry of a factor i.i.d. from `probs` and rejects a triple iff u, v or w is all zero
 * (utils.py:222-232).  The three factors of an accepted triple are therefore independent, each distributed as an i.i.d.
 * factor CONDITIONED on being non-zero -- which can be sampled directly, without a rejection loop: a factor is cut into
 * groups of three consecutive entries (the last group holds what is left: S = 4 -> 3+1, 9 -> 3+3+3, 16 -> 3*5+1), groups
 * are drawn left to right, and while every group so far is all zero the next group is drawn from the distribution tilted
 * by the probability that the REST of the factor can still make it non-zero:
 *     P(x_g | all earlier groups zero) ~ P(x_g) * (x_g != 0 ? 1 : 1 - P(all later groups zero)),
 * otherwise from the plain product distribution.  Each group distribution (<= 5^3 = 125 outcomes) is an alias table of
 * 128 buckets (Vose's construction in the order written below), a bucket being a 9-bit threshold and an alias outcome:
 * one 16-bit draw h picks bucket h & 127 and keeps the bucket's own outcome iff (h >> 7) < thr, else takes the alias.
 * Outcome o of a size-3 group encodes the value indices (o % n, o / n % n, o / n^2).
 * Draw m = f * NG + g (factor f, group g, NG groups per factor) is half m & 1 of word (m >> 1) & 3 of Philox block m >> 3
 * with ctr = (d_lo, d_hi, r, block), key = seed.  No tries: max_tries does not apply.  Used when n_values <= 5 and
 * P(0) <= 0.999 (orc_alias_applies); other alphabets keep the thresholded contract above. */
#define ORC_AB 128
typedef struct {
    uint16_t tab[8][ORC_AB]; /* [0] plain size 3, [1] plain size 1, [2 + g] tilted table of group g */
    int32_t ng, last, n, zo3, zo1;
} orc_alias;

static void orc_vose(const double *P, int count, uint16_t *out) {
    double sc[ORC_AB], prob[ORC_AB];
    int alias[ORC_AB], small[ORC_AB], large[ORC_AB], ns = 0, nl = 0;
    for (int o = 0; o < ORC_AB; o++) {
        sc[o] = (o < count ? P[o] : 0.0) * (double)ORC_AB;
        alias[o] = o;
        prob[o] = 1.0;
    }
    for (int o = 0; o < ORC_AB; o++) {
        if (sc[o] < 1.0) small[ns++] = o; else large[nl++] = o;
    }
    while (ns > 0 && nl > 0) {
        const int sm = small[--ns], lg = large[--nl];
        prob[sm] = sc[sm];
        alias[sm] = lg;
        sc[lg] = (sc[lg] + sc[sm]) - 1.0;
        if (sc[lg] < 1.0) small[ns++] = lg; else large[nl++] = lg;
    }
    for (int o = 0; o < ORC_AB; o++) {
        double q = prob[o] < 0.0 ? 0.0 : (prob[o] > 1.0 ? 1.0 : prob[o]);
        uint32_t thr = (uint32_t)(q * 512.0 + 0.5);
        int al = alias[o];
        if (thr >= 512u) thr = 511u, al = o; /* probability one: either branch gives the bucket's own outcome */
        out[o] = (uint16_t)(thr | ((uint32_t)al << 9));
    }
}

int orc_alias_applies(const int32_t *values, const double *probs, int n_values, int S) {
    if (n_values < 1 || n_values > 5 || !(S == 4 || S == 9 || S == 16)) return 0;
    double total = 0, p0 = 0;
    for (int i = 0; i < n_values; i++) total += probs[i];
    for (int i = 0; i < n_values; i++)
        if (values[i] == 0) p0 += probs[i] / total;
    return p0 <= 0.999;
}

static void orc_alias_build(const int32_t *values, const double *probs, int n, int S, orc_alias *A) {
    double p[5], total = 0, p0 = 0;
    int z = -1;
    for (int i = 0; i < n; i++) total += probs[i];
    for (int i = 0; i < n; i++) {
        p[i] = probs[i] / total;
        if (values[i] == 0 && z < 0) z = i, p0 = p[i];
    }
    memset(A, 0, sizeof(*A));
    A->n = n, A->ng = (S + 2) / 3, A->last = S - 3 * (A->ng - 1);
    A->zo3 = z >= 0 ? z + n * z + n * n * z : 255;
    A->zo1 = z >= 0 ? z : 255;
    double P3[ORC_AB], P1[ORC_AB];
    for (int o = 0; o < n * n * n; o++) P3[o] = (p[o % n] * p[(o / n) % n]) * p[o / (n * n)];
    for (int o = 0; o < n; o++) P1[o] = p[o];
    orc_vose(P3, n * n * n, A->tab[0]);
    orc_vose(P1, n, A->tab[1]);
    for (int g = 0; g < A->ng; g++) {
        const int size = (g < A->ng - 1) ? 3 : A->last;
        const int count = size == 3 ? n * n * n : n, zo = size == 3 ? A->zo3 : A->zo1;
        double qlater = 1.0; /* P(every entry after this group is zero) */
        for (int e = 3 * g + size; e < S; e++) qlater *= p0;
        double W[ORC_AB], sum = 0;
        for (int o = 0; o < count; o++) {
            W[o] = size == 3 ? P3[o] : P1[o];
            if (o == zo) W[o] *= (1.0 - qlater);
            sum += W[o];
        }
        for (int o = 0; o < count; o++) W[o] /= sum;
        orc_vose(W, count, A->tab[2 + g]);
    }
}

/* tables as the library's tg_demo_alias_tables writes them: uint16 [8][128] */
void orc_alias_tables(const int32_t *values, const double *probs, int n_values, int S, uint16_t *out) {
    orc_alias A;
    orc_alias_build(values, probs, n_values, S, &A);
    memcpy(out, A.tab, sizeof(A.tab));
}

static void orc_demo_philox_v2(uint64_t seed, uint64_t d, const int32_t *values, const orc_alias *A, int R, int S, int shift,
                               int32_t *tokens_out, int32_t *target_out) {
    const int S3 = S * S * S, n = A->n;
    int32_t tmp[4096];
    memset(target_out, 0, sizeof(int32_t) * S3);
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < R; r++) {
        int32_t f[3][64];
        uint32_t blk[4] = {0, 0, 0, 0};
        int az = 1; /* all groups of the current factor so far are zero */
        for (int m = 0; m < 3 * A->ng; m++) {
            if ((m & 7) == 0) {
                uint32_t ctr[4] = {(uint32_t)d, (uint32_t)(d >> 32), (uint32_t)r, (uint32_t)(m >> 3)};
                orc_philox4x32_10(ctr, key, blk);
            }
            const int fac = m / A->ng, g = m % A->ng;
            if (g == 0) az = 1;
            const uint32_t word = blk[(m >> 1) & 3], h = (m & 1) ? (word >> 16) : (word & 0xFFFFu);
            const int size = (g < A->ng - 1) ? 3 : A->last;
            const uint16_t b = A->tab[az ? 2 + g : (size == 3 ? 0 : 1)][h & 127];
            const int o = ((h >> 7) < (uint32_t)(b & 511)) ? (int)(h & 127) : (int)(b >> 9);
            if (size == 3) {
                f[fac][3 * g] = values[o % n], f[fac][3 * g + 1] = values[(o / n) % n], f[fac][3 * g + 2] = values[o / (n * n)];
                az = az && o == A->zo3;
            } else {
                f[fac][3 * g] = values[o];
                az = az && o == A->zo1;
            }
        }
        int32_t *tok = tokens_out + r * 3 * S;
        for (int m = 0; m < 3; m++)
            for (int i = 0; i < S; i++) tok[m * S + i] = f[m][i] + shift;
        orc_uvw_to_tensor(f[0], f[1], f[2], S, tmp);
        for (int e = 0; e < S3; e++) target_out[e] += tmp[e];
    }
}

void orc_demos_philox_v2_batch(uint64_t seed, uint64_t d0, int64_t n, const int32_t *values, const double *probs, int n_values,
                               int R, int S, int shift, int32_t *tokens_out, int32_t *targets_out) {
    const int64_t S3 = (int64_t)S * S * S;
    orc_alias A;
    orc_alias_build(values, probs, n_values, S, &A);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
        orc_demo_philox_v2(seed, d0 + (uint64_t)i, values, &A, R, S, shift, tokens_out + i * R * 3 * S, targets_out + i * S3);
}

/* Random unimodular triple (ours): M_f = L*U for f = 0,1,2 (A,B,C).  Entry
 * (r,c) of the draw grid uses byte (r*S+c)%16 (little-endian within the four
 * words) of Philox block (r*S+c)/16 with ctr = (block, f, 0x6D617473, d_lo)
 * and the demo-stream key; low 7 bits < thr_nz => magnitude 1, top bit => sign.
 * L takes the draws below the diagonal (unit diagonal), U those above it and
 * the sign of the diagonal draw as its +-1 diagonal.  mats [3][S][S]. */
void orc_sample_unimodular(uint64_t seed, uint64_t d, int S, uint32_t thr_nz, int32_t *mats) {
    uint32_t key[2] = {(uint32_t)seed ^ ((uint32_t)(d >> 32) * 0x9E3779B9u), (uint32_t)(seed >> 32)};
    for (int f = 0; f < 3; f++) {
        int32_t L[256], U[256];
        uint32_t blk[4] = {0, 0, 0, 0};
        for (int e = 0; e < S * S; e++) {
            if ((e & 15) == 0) {
                uint32_t ctr[4] = {(uint32_t)(e >> 4), (uint32_t)f, 0x6D617473u, (uint32_t)d};
                orc_philox4x32_10(ctr, key, blk);
            }
            uint32_t byte = (blk[(e >> 2) & 3] >> (8 * (e & 3))) & 0xFFu;
            int r = e / S, c = e % S;
            int mag = (byte & 0x7Fu) < thr_nz ? 1 : 0;
            int val = (byte & 0x80u) ? -mag : mag;
            L[e] = r > c ? val : (r == c ? 1 : 0);
            U[e] = r < c ? val : (r == c ? ((byte & 0x80u) ? -1 : 1) : 0);
        }
        for (int r = 0; r < S; r++)
            for (int c = 0; c < S; c++) {
                int32_t acc = 0;
                for (int k = 0; k < S; k++) acc += L[r * S + k] * U[k * S + c];
                mats[(f * S + r) * S + c] = acc;
            }
    }
}

/* ------------------------------------------------------------------ */
/* state key (ours; replaces the string key of utils.py:164-169)       */
/* ------------------------------------------------------------------ */
/* 64-bit key of a head residual: sum over entries e of mix(e, value) where
 * mix is the splitmix64 finaliser of (value + 2^32 * (e+1)); entries equal to
 * zero contribute nothing, so the all-zero state has key 0.  Order-free sum
 * => reducible in parallel.  Equal states <=> equal strings in the reference;
 * equal keys are equal states up to a 2^-64 collision. */
static uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* key(T) = sum_{i,j,k} T[i][j][k] * A_i * B_j * C_k  (mod 2^64),  A_i = splitmix64(0x1000 + i) | 1,
 * B_j = splitmix64(0x2000 + j) | 1,  C_k = splitmix64(0x3000 + k) | 1: the trilinear form of T at three fixed odd 64-bit
 * vectors (distinct per mode, so transposed states get different keys).  Linear, so the key of a child state is the key
 * of its parent minus the contribution of the rank-1 action -- which is just (sum u_i A_i)(sum v_j B_j)(sum w_k C_k).
 * The all-zero state has key 0. */
uint64_t orc_state_key(const int32_t *T, int S) {
    uint64_t h = 0;
    for (int i = 0; i < S; i++) {
        const uint64_t a = splitmix64(0x1000ull + (uint64_t)i) | 1ull;
        for (int j = 0; j < S; j++) {
            const uint64_t ab = a * (splitmix64(0x2000ull + (uint64_t)j) | 1ull);
            for (int k = 0; k < S; k++) {
                const int32_t t = T[(i * S + j) * S + k];
                if (t != 0) h += (uint64_t)(int64_t)t * (ab * (splitmix64(0x3000ull + (uint64_t)k) | 1ull));
            }
        }
    }
    return h;
}

void orc_state_key_batch(const int32_t *T, int64_t B, int S, uint64_t *keys) {
    const int64_t S3 = (int64_t)S * S * S;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; b++) keys[b] = orc_state_key(T + b * S3, S);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* all the host threads the process may use, whatever OMP_NUM_THREADS said at start-up (torchrun exports
 * OMP_NUM_THREADS=1 to its workers); returns the new thread count */
int orc_use_all_threads(void) {
#ifdef _OPENMP
    omp_set_num_threads(omp_get_num_procs());
    return omp_get_max_threads();
#else
    return 1;
#endif
}
