// tg_common.cuh -- shared device helpers for the TensorGame kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tensorgame.h"

namespace tg {

// ---------------------------------------------------------------- geometry
// Device slab: int8 [B][GP]; entry (i,j,k) at i*RP + j*S + k.  One 32-bit word
// holds four consecutive (j,k) entries of one i-row, so the rank-1 update of a
// word is  u_i * pack(v_j w_k)  -- one IMAD per word (see tg_step.cu).
template <int S>
struct Geo {
    static constexpr int S2 = S * S;
    static constexpr int S3 = S2 * S;
    static constexpr int RP = (S2 + 3) & ~3;       // row pitch (bytes)
    static constexpr int WR = RP / 4;              // 32-bit words per row
    static constexpr int GP = (S * RP + 15) & ~15; // game pitch (bytes)
    static constexpr int TP = (3 * S + 15) & ~15;  // token pitch (bytes)
    static constexpr bool STRADDLE = (S % 4) != 0; // a word can span two j
};

__host__ __device__ constexpr bool supported_S(int S) { return S == 4 || S == 9 || S == 16; }

extern int g_last_cuda_error;

// Kernel variants for tuning sweeps exist only in the -DTG_TUNING build (libtensorgame_b200_tuning.so, selected with
// TG_TUNING=1 by mat_mul_b200/_lib.py): there an environment variable picks the variant; the production library compiles
// the measured-best path alone and reads no environment.
#ifdef TG_TUNING
}
#include <cstdlib>
namespace tg {
inline int tuning_env(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
#endif

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return TG_E_CUDA;
}
#define TG_CUDA(call)                                  \
    do {                                               \
        cudaError_t _e = (call);                       \
        if (_e != cudaSuccess) return tg::cuda_fail(_e); \
    } while (0)

// change of basis: set by a fast kernel on the games it leaves to the exact int32 kernel, which clears it
constexpr uint8_t BASIS_REDO = 0x80;
// tg_basis_mma.cu: 16x16x16 games, one warp per game on mma.sync int8 + f16; tg_basis_mma9.cu: 9x9x9 games on mma.sync f16.
// slab_out is an int8 slab (out16 == 0) or an int16 slab.
int launch_basis_mma16(const int8_t *slab_in, const int8_t *mats, long long mat_stride, void *slab_out, int out16, uint8_t *flags,
                       long long N, cudaStream_t st);
int launch_basis_mma9(const int8_t *slab_in, const int8_t *mats, long long mat_stride, void *slab_out, int out16, uint8_t *flags,
                      long long N, cudaStream_t st);

// tg_rollout9.cu: tg_rollout / tg_replay at 9x9x9 and 16x16x16, one thread per row of a game
int launch_rollout_rows(const int8_t *slab_in, const uint8_t *tape, long long stride, int K, int8_t *slab_out, uint8_t *flags,
                        int32_t *nnz, int32_t *steps, long long B, int S, int shift, int freeze, cudaStream_t st);
// tg_demo_mma.cu: sum of the R rank-1 terms of 16x16x16 action lists, one warp per demo on mma.sync f16 (R <= 64)
bool demo_acc16_mma_applies(int R);
int launch_demo_acc16_mma(const uint8_t *tape, long long tape_step_stride, long long N, int R, int shift, int8_t *slab,
                          uint8_t *flags, int or_flags, cudaStream_t st);

// ---------------------------------------------------------------- PTX: mbarrier + bulk async copy (TMA 1-D)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted on the mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by bulk async-groups
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (bulk store source)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- SWAR helpers on 4 packed int8
constexpr uint32_t H4 = 0x80808080u;
constexpr uint32_t ONES4 = 0x01010101u;

// bit 7 of each byte set iff that byte is non-zero
__device__ __forceinline__ uint32_t nonzero_mask(uint32_t x) { return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & H4; }

// sum of the four unsigned bytes of x
__device__ __forceinline__ uint32_t byte_sum(uint32_t x) { return __dp4a(x, ONES4, 0u); }

// vector accesses as PTX, so that the access width is exactly the one the alignment test in front of it allows (the
// compiler re-vectorised the plain C++ stores of the misaligned branch into an 8-byte store at p + 2)
__device__ __forceinline__ void st_f32x4(float *p, float a, float b, float c, float d) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(p)), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_f32x2(float *p, float a, float b) {
    asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(p)), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_f32(float *p, float a) {
    asm volatile("st.global.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(a) : "memory");
}
__device__ __forceinline__ void ld_f32x4(const float *p, float f[4]) {
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "l"(__cvta_generic_to_global(p)));
}
__device__ __forceinline__ void ld_f32x2(const float *p, float &a, float &b) {
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "l"(__cvta_generic_to_global(p)));
}
__device__ __forceinline__ float ld_f32(const float *p) {
    float a;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(a) : "l"(__cvta_generic_to_global(p)));
    return a;
}

// four consecutive float32 entries of one output row (entries 4c .. 4c+3, the first nv of them inside the row) leave with
// the widest stores the address allows: one 16-byte store when S^2 is a multiple of four (4x4x4, 16x16x16: every run of a
// 16-byte aligned batch is aligned), at 9x9x9 (rows of 81 floats: the alignment of a run alternates with the row) two
// 8-byte stores, or 4 + 8 + 4 bytes around the aligned pair in the middle.  p points to global memory.
template <int S>
__device__ __forceinline__ void store_run(float *p, int nv, float a, float b, float c, float d) {
    const uint32_t lo = (uint32_t)reinterpret_cast<uintptr_t>(p);
    if constexpr (Geo<S>::S2 % 4 == 0) {
        if ((lo & 15u) == 0) {
            st_f32x4(p, a, b, c, d);
            return;
        }
    }
    if (nv == 4) {
        if ((lo & 7u) == 0) {
            st_f32x2(p, a, b);
            st_f32x2(p + 2, c, d);
        } else {
            st_f32(p, a);
            st_f32x2(p + 1, b, c);
            st_f32(p + 3, d);
        }
    } else {
        if (nv > 0) st_f32(p, a);
        if (nv > 1) st_f32(p + 1, b);
        if (nv > 2) st_f32(p + 2, c);
    }
}

// counterpart of store_run: four consecutive float32 entries of a row (the first nv of them exist) with the widest loads
// the address allows
template <int S>
__device__ __forceinline__ void load_run(const float *p, int nv, float f[4]) {
    const uint32_t lo = (uint32_t)reinterpret_cast<uintptr_t>(p);
    f[0] = f[1] = f[2] = f[3] = 0.f;
    if constexpr (Geo<S>::S2 % 4 == 0) {
        if ((lo & 15u) == 0) {
            ld_f32x4(p, f);
            return;
        }
    }
    if (nv == 4) {
        if ((lo & 7u) == 0) {
            ld_f32x2(p, f[0], f[1]);
            ld_f32x2(p + 2, f[2], f[3]);
        } else {
            f[0] = ld_f32(p);
            ld_f32x2(p + 1, f[1], f[2]);
            f[3] = ld_f32(p + 3);
        }
    } else {
        if (nv > 0) f[0] = ld_f32(p);
        if (nv > 1) f[1] = ld_f32(p + 1);
        if (nv > 2) f[2] = ld_f32(p + 2);
    }
}

} // namespace tg
