// tg_host.cu -- host-buffer entry points: the end-to-end path a caller with
// data in host memory takes (H2D copy -> kernel -> D2H copy).  Chunks flow through three dedicated streams -- every H2D
// copy on one, every kernel on another, every D2H copy on a third -- chained by events per staging buffer, so that the
// two copy engines stream back to back: a chunk's upload never queues behind another chunk's download (with one stream
// per buffer it did: at 8 ranks per host the path reached half of the raw pinned-copy rate).
#include <new>

#include "tg_common.cuh"

struct tg_host_ctx {
    int device;
    int S;
    int64_t chunk;
    static constexpr int NBUF = 4;
    cudaStream_t s_in, s_k, s_out;
    cudaEvent_t ev_in[NBUF], ev_k[NBUF], ev_out[NBUF];
    int8_t *slab[NBUF];
    uint8_t *tape[NBUF];
    uint8_t *flags[NBUF];
    int32_t *nnz[NBUF];
    int32_t *steps[NBUF];
    uint8_t *tape_k[NBUF]; // multi-step tape staging [K][chunk][TP], grown on demand
    int64_t tape_k_bytes[NBUF];
};

// device staging for a K-step tape of one chunk (kept in the context, grown when a call needs more)
static int ensure_tape_k(tg_host_ctx *c, int i, int64_t bytes) {
    if (c->tape_k_bytes[i] >= bytes) return TG_OK;
    TG_CUDA(cudaDeviceSynchronize());
    cudaFree(c->tape_k[i]);
    c->tape_k[i] = nullptr, c->tape_k_bytes[i] = 0;
    TG_CUDA(cudaMalloc(&c->tape_k[i], (size_t)bytes));
    c->tape_k_bytes[i] = bytes;
    return TG_OK;
}

extern "C" {

int tg_host_ctx_create(tg_host_ctx **out, int device, int S, int64_t max_chunk) {
    if (!out || !tg::supported_S(S) || max_chunk <= 0) return TG_E_ARG;
    int rp, gp, tp;
    tg_layout(S, &rp, &gp, &tp);
    TG_CUDA(cudaSetDevice(device));
    tg_host_ctx *c = new (std::nothrow) tg_host_ctx();
    if (!c) return TG_E_ARG;
    c->device = device, c->S = S, c->chunk = max_chunk;
    c->s_in = c->s_k = c->s_out = nullptr;
    for (int i = 0; i < tg_host_ctx::NBUF; i++) {
        c->ev_in[i] = c->ev_k[i] = c->ev_out[i] = nullptr;
        c->slab[i] = nullptr, c->tape[i] = nullptr, c->flags[i] = nullptr, c->nnz[i] = nullptr;
        c->steps[i] = nullptr, c->tape_k[i] = nullptr, c->tape_k_bytes[i] = 0;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    for (int i = 0; i < tg_host_ctx::NBUF && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&c->slab[i], (size_t)max_chunk * gp);
        if (e == cudaSuccess) e = cudaMalloc(&c->tape[i], (size_t)max_chunk * tp);
        if (e == cudaSuccess) e = cudaMalloc(&c->flags[i], (size_t)max_chunk);
        if (e == cudaSuccess) e = cudaMalloc(&c->nnz[i], (size_t)max_chunk * 4);
        if (e == cudaSuccess) e = cudaMalloc(&c->steps[i], (size_t)max_chunk * 4);
    }
    if (e != cudaSuccess) {
        tg_host_ctx_destroy(c);
        return tg::cuda_fail(e);
    }
    *out = c;
    return TG_OK;
}

int tg_host_ctx_destroy(tg_host_ctx *c) {
    if (!c) return TG_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < tg_host_ctx::NBUF; i++) {
        cudaFree(c->slab[i]);
        cudaFree(c->tape[i]);
        cudaFree(c->flags[i]);
        cudaFree(c->nnz[i]);
        cudaFree(c->steps[i]);
        cudaFree(c->tape_k[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    }
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_k) cudaStreamDestroy(c->s_k);
    if (c->s_out) cudaStreamDestroy(c->s_out);
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
        if (q >= tg_host_ctx::NBUF) TG_CUDA(cudaStreamWaitEvent(c->s_in, c->ev_out[i], 0)); // the buffer's last download is done
        TG_CUDA(cudaMemcpyAsync(c->slab[i], slab_in + b0 * gp, (size_t)n * gp, cudaMemcpyHostToDevice, c->s_in));
        TG_CUDA(cudaMemcpyAsync(c->tape[i], tape + b0 * tp, (size_t)n * tp, cudaMemcpyHostToDevice, c->s_in));
        TG_CUDA(cudaEventRecord(c->ev_in[i], c->s_in));
        TG_CUDA(cudaStreamWaitEvent(c->s_k, c->ev_in[i], 0));
        int rc = tg_step(c->slab[i], c->tape[i], c->slab[i], c->flags[i], c->nnz[i], n, c->S, shift, c->s_k);
        if (rc != TG_OK) return rc;
        TG_CUDA(cudaEventRecord(c->ev_k[i], c->s_k));
        TG_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[i], 0));
        TG_CUDA(cudaMemcpyAsync(slab_out + b0 * gp, c->slab[i], (size_t)n * gp, cudaMemcpyDeviceToHost, c->s_out));
        TG_CUDA(cudaMemcpyAsync(flags + b0, c->flags[i], (size_t)n, cudaMemcpyDeviceToHost, c->s_out));
        TG_CUDA(cudaMemcpyAsync(nnz + b0, c->nnz[i], (size_t)n * 4, cudaMemcpyDeviceToHost, c->s_out));
        TG_CUDA(cudaEventRecord(c->ev_out[i], c->s_out));
    }
    TG_CUDA(cudaStreamSynchronize(c->s_out));
    return TG_OK;
}

int tg_rollout_host(tg_host_ctx *c, const int8_t *slab_in, const uint8_t *tape, int K, int8_t *slab_out, uint8_t *flags,
                    int32_t *nnz, int32_t *steps, int64_t B, int shift) {
    if (!c || B < 0 || K < 0) return TG_E_ARG;
    if (B == 0) return TG_OK;
    if (!slab_in || (K > 0 && !tape) || !slab_out || !flags || !nnz || !steps) return TG_E_ARG;
    int rp, gp, tp;
    tg_layout(c->S, &rp, &gp, &tp);
    TG_CUDA(cudaSetDevice(c->device));
    for (int i = 0; i < tg_host_ctx::NBUF; i++) {
        const int rc = ensure_tape_k(c, i, (int64_t)(K > 0 ? K : 1) * c->chunk * tp);
        if (rc != TG_OK) return rc;
    }
    int q = 0;
    for (int64_t b0 = 0; b0 < B; b0 += c->chunk, q++) {
        const int i = q % tg_host_ctx::NBUF;
        const int64_t n = (B - b0 < c->chunk) ? B - b0 : c->chunk;
        if (q >= tg_host_ctx::NBUF) TG_CUDA(cudaStreamWaitEvent(c->s_in, c->ev_out[i], 0));
        TG_CUDA(cudaMemcpyAsync(c->slab[i], slab_in + b0 * gp, (size_t)n * gp, cudaMemcpyHostToDevice, c->s_in));
        // the chunk's columns of the step-major host tape [K][B][TP] -> a dense [K][n][TP] staging tape
        if (K > 0)
            TG_CUDA(cudaMemcpy2DAsync(c->tape_k[i], (size_t)n * tp, tape + b0 * tp, (size_t)B * tp, (size_t)n * tp, (size_t)K,
                                      cudaMemcpyHostToDevice, c->s_in));
        TG_CUDA(cudaEventRecord(c->ev_in[i], c->s_in));
        TG_CUDA(cudaStreamWaitEvent(c->s_k, c->ev_in[i], 0));
        int rc = tg_rollout(c->slab[i], c->tape_k[i], n * tp, K, c->slab[i], c->flags[i], c->nnz[i], c->steps[i], n, c->S, shift, c->s_k);
        if (rc != TG_OK) return rc;
        TG_CUDA(cudaEventRecord(c->ev_k[i], c->s_k));
        TG_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[i], 0));
        TG_CUDA(cudaMemcpyAsync(slab_out + b0 * gp, c->slab[i], (size_t)n * gp, cudaMemcpyDeviceToHost, c->s_out));
        TG_CUDA(cudaMemcpyAsync(flags + b0, c->flags[i], (size_t)n, cudaMemcpyDeviceToHost, c->s_out));
        TG_CUDA(cudaMemcpyAsync(nnz + b0, c->nnz[i], (size_t)n * 4, cudaMemcpyDeviceToHost, c->s_out));
        TG_CUDA(cudaMemcpyAsync(steps + b0, c->steps[i], (size_t)n * 4, cudaMemcpyDeviceToHost, c->s_out));
        TG_CUDA(cudaEventRecord(c->ev_out[i], c->s_out));
    }
    TG_CUDA(cudaStreamSynchronize(c->s_out));
    return TG_OK;
}

int tg_demo_gen_host(tg_host_ctx *c, uint64_t seed, uint64_t first_demo, int64_t N, int R, int shift, const int8_t *values,
                     const double *probs, int n_values, int max_tries, uint8_t *tape, int8_t *slab, uint8_t *flags) {
    if (!c || N < 0 || R < 1) return TG_E_ARG;
    if (N == 0) return TG_OK;
    if (!tape || !slab || !flags) return TG_E_ARG;
    int rp, gp, tp;
    tg_layout(c->S, &rp, &gp, &tp);
    TG_CUDA(cudaSetDevice(c->device));
    for (int i = 0; i < tg_host_ctx::NBUF; i++) {
        const int rc = ensure_tape_k(c, i, (int64_t)R * c->chunk * tp);
        if (rc != TG_OK) return rc;
    }
    int q = 0;
    for (int64_t n0 = 0; n0 < N; n0 += c->chunk, q++) {
        const int i = q % tg_host_ctx::NBUF;
        const int64_t n = (N - n0 < c->chunk) ? N - n0 : c->chunk;
        if (q >= tg_host_ctx::NBUF) TG_CUDA(cudaStreamWaitEvent(c->s_k, c->ev_out[i], 0)); // the buffer's last download is done
        int rc = tg_demo_gen_philox(seed, first_demo + (uint64_t)n0, n, R, c->S, shift, values, probs, n_values, max_tries,
                                    c->tape_k[i], n * tp, c->slab[i], c->flags[i], c->s_k);
        if (rc != TG_OK) return rc;
        TG_CUDA(cudaEventRecord(c->ev_k[i], c->s_k));
        TG_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[i], 0));
        // dense [R][n][TP] staging tape -> the chunk's columns of the step-major host tape [R][N][TP]
        TG_CUDA(cudaMemcpy2DAsync(tape + n0 * tp, (size_t)N * tp, c->tape_k[i], (size_t)n * tp, (size_t)n * tp, (size_t)R,
                                  cudaMemcpyDeviceToHost, c->s_out));
        TG_CUDA(cudaMemcpyAsync(slab + n0 * gp, c->slab[i], (size_t)n * gp, cudaMemcpyDeviceToHost, c->s_out));
        TG_CUDA(cudaMemcpyAsync(flags + n0, c->flags[i], (size_t)n, cudaMemcpyDeviceToHost, c->s_out));
        TG_CUDA(cudaEventRecord(c->ev_out[i], c->s_out));
    }
    TG_CUDA(cudaStreamSynchronize(c->s_out));
    return TG_OK;
}

} // extern "C"
