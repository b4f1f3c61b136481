// tg_host.cu -- host-buffer entry points: the end-to-end path a caller with
// data in host memory takes (H2D copy -> kernel -> D2H copy, chunked and
// pipelined over several streams so that PCIe in, compute and PCIe out overlap).
#include <new>

#include "tg_common.cuh"

struct tg_host_ctx {
    int device;
    int S;
    int64_t chunk;
    static constexpr int NBUF = 3;
    cudaStream_t stream[NBUF];
    int8_t *slab[NBUF];
    uint8_t *tape[NBUF];
    uint8_t *flags[NBUF];
    int32_t *nnz[NBUF];
};

extern "C" {

int tg_host_ctx_create(tg_host_ctx **out, int device, int S, int64_t max_chunk) {
    if (!out || !tg::supported_S(S) || max_chunk <= 0) return TG_E_ARG;
    int rp, gp, tp;
    tg_layout(S, &rp, &gp, &tp);
    TG_CUDA(cudaSetDevice(device));
    tg_host_ctx *c = new (std::nothrow) tg_host_ctx();
    if (!c) return TG_E_ARG;
    c->device = device, c->S = S, c->chunk = max_chunk;
    for (int i = 0; i < tg_host_ctx::NBUF; i++) {
        c->stream[i] = nullptr, c->slab[i] = nullptr, c->tape[i] = nullptr, c->flags[i] = nullptr, c->nnz[i] = nullptr;
    }
    for (int i = 0; i < tg_host_ctx::NBUF; i++) {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&c->slab[i], (size_t)max_chunk * gp);
        if (e == cudaSuccess) e = cudaMalloc(&c->tape[i], (size_t)max_chunk * tp);
        if (e == cudaSuccess) e = cudaMalloc(&c->flags[i], (size_t)max_chunk);
        if (e == cudaSuccess) e = cudaMalloc(&c->nnz[i], (size_t)max_chunk * 4);
        if (e != cudaSuccess) {
            tg_host_ctx_destroy(c);
            return tg::cuda_fail(e);
        }
    }
    *out = c;
    return TG_OK;
}

int tg_host_ctx_destroy(tg_host_ctx *c) {
    if (!c) return TG_OK;
    cudaSetDevice(c->device);
    for (int i = 0; i < tg_host_ctx::NBUF; i++) {
        if (c->stream[i]) cudaStreamSynchronize(c->stream[i]);
        cudaFree(c->slab[i]);
        cudaFree(c->tape[i]);
        cudaFree(c->flags[i]);
        cudaFree(c->nnz[i]);
        if (c->stream[i]) cudaStreamDestroy(c->stream[i]);
    }
    delete c;
    return TG_OK;
}

int tg_step_host(tg_host_ctx *c, const int8_t *slab_in, const uint8_t *tape, int8_t *slab_out, uint8_t *flags,
                 int32_t *nnz, int64_t B, int shift) {
    if (!c || B < 0) return TG_E_ARG;
    if (B == 0) return TG_OK;
    if (!slab_in || !tape || !slab_out || !flags || !nnz) return TG_E_ARG;
    int rp, gp, tp;
    tg_layout(c->S, &rp, &gp, &tp);
    TG_CUDA(cudaSetDevice(c->device));
    int q = 0;
    for (int64_t b0 = 0; b0 < B; b0 += c->chunk, q++) {
        const int i = q % tg_host_ctx::NBUF;
        const int64_t n = (B - b0 < c->chunk) ? B - b0 : c->chunk;
        cudaStream_t st = c->stream[i];
        TG_CUDA(cudaMemcpyAsync(c->slab[i], slab_in + b0 * gp, (size_t)n * gp, cudaMemcpyHostToDevice, st));
        TG_CUDA(cudaMemcpyAsync(c->tape[i], tape + b0 * tp, (size_t)n * tp, cudaMemcpyHostToDevice, st));
        int rc = tg_step(c->slab[i], c->tape[i], c->slab[i], c->flags[i], c->nnz[i], n, c->S, shift, st);
        if (rc != TG_OK) return rc;
        TG_CUDA(cudaMemcpyAsync(slab_out + b0 * gp, c->slab[i], (size_t)n * gp, cudaMemcpyDeviceToHost, st));
        TG_CUDA(cudaMemcpyAsync(flags + b0, c->flags[i], (size_t)n, cudaMemcpyDeviceToHost, st));
        TG_CUDA(cudaMemcpyAsync(nnz + b0, c->nnz[i], (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < tg_host_ctx::NBUF; i++) TG_CUDA(cudaStreamSynchronize(c->stream[i]));
    return TG_OK;
}

} // extern "C"
