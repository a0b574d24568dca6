// Host <-> device transfers for the host-pointer drop-ins.
//
// The reference's callers hand over ordinary Rust `Vec`s, i.e. PAGEABLE memory.  cudaMemcpyAsync from pageable memory
// is staged by the driver through one internal buffer with a single CPU thread: 512 MiB of scalars take ~55 ms, longer
// than the whole 2^24 MSM.  Here a few host threads copy pieces of the caller's buffer into their own pinned slots and
// enqueue the DMA of each piece on their own stream, so that the CPU copy of piece p + 1 overlaps the DMA of piece p
// and several cores share the memory-bound copy.  Pinned / registered caller memory skips all of this.
#include <thread>

#include "common.h"

namespace h2b {

static const int STAGE_THREADS_MAX = 8;
static const int STAGE_SLOTS = 2;
// 4 MiB pieces; transfers below two pieces go through the driver's path (spawning the copy threads costs ~0.3 ms).
// H2B_STAGE_PIECE_LOG shrinks the piece (tests).
static size_t stage_piece() {
    static size_t v = 0;
    if (!v) {
        const char* e = getenv("H2B_STAGE_PIECE_LOG");
        int lg = e ? atoi(e) : 22;
        if (lg < 8 || lg > 26) lg = 22;
        v = (size_t)1 << lg;
    }
    return v;
}
#define STAGE_PIECE (stage_piece())
#define STAGE_MIN_BYTES (stage_piece() * 2)

struct Stager {
    int nthreads = 0;
    void* slot[STAGE_THREADS_MAX][STAGE_SLOTS];
    cudaStream_t stream[STAGE_THREADS_MAX];
    cudaEvent_t slot_free[STAGE_THREADS_MAX][STAGE_SLOTS];
    cudaEvent_t done[STAGE_THREADS_MAX];
    cudaEvent_t ready = nullptr;
};

static int stager_init(DeviceCtx& ctx) {
    if (ctx.stager) return H2B_OK;
    Stager* s = new Stager();
    unsigned hw = std::thread::hardware_concurrency();
    int want = hw >= 16 ? 8 : (hw >= 8 ? 4 : 2);
    // several devices of one process stage at the same time (point-range split, round-robin columns): share the cores
    const int nd = h2b_device_count();
    if (nd > 1 && hw) { int share = (int)hw / nd; if (share < 2) share = 2; if (want > share) want = share; }
    const char* e = getenv("H2B_STAGE_THREADS");
    if (e && atoi(e) >= 1 && atoi(e) <= STAGE_THREADS_MAX) want = atoi(e);
    s->nthreads = want;
    for (int t = 0; t < s->nthreads; ++t) {
        H2B_CUDA(cudaStreamCreateWithFlags(&s->stream[t], cudaStreamNonBlocking));
        H2B_CUDA(cudaEventCreateWithFlags(&s->done[t], cudaEventDisableTiming));
        for (int k = 0; k < STAGE_SLOTS; ++k) {
            H2B_CUDA(cudaMallocHost(&s->slot[t][k], STAGE_PIECE));
            H2B_CUDA(cudaEventCreateWithFlags(&s->slot_free[t][k], cudaEventDisableTiming));
        }
    }
    H2B_CUDA(cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming));
    ctx.stager = s;
    return H2B_OK;
}

void stager_release(DeviceCtx& ctx) {
    Stager* s = ctx.stager;
    if (!s) return;
    for (int t = 0; t < s->nthreads; ++t) {
        cudaStreamSynchronize(s->stream[t]);
        for (int k = 0; k < STAGE_SLOTS; ++k) { cudaFreeHost(s->slot[t][k]); cudaEventDestroy(s->slot_free[t][k]); }
        cudaEventDestroy(s->done[t]);
        cudaStreamDestroy(s->stream[t]);
    }
    cudaEventDestroy(s->ready);
    delete s;
    ctx.stager = nullptr;
}

bool host_is_pageable(const void* p) {
#ifdef H2B_EMU
    (void)p;
    return true;
#else
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
#endif
}

// After the call returns, everything enqueued on `consumer` afterwards sees the data in d_dst.  For pageable sources
// the host-side copies are complete on return (the caller may reuse h_src); the DMA may still be in flight.
// order_after_consumer: the DMAs wait for the work already queued on `consumer` (needed when that work may still read
// d_dst); pass false when d_dst is not touched by anything in flight, so that the upload overlaps the consumer's work.
int host_upload(DeviceCtx& ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t consumer, bool order_after_consumer) {
    if (bytes == 0) return H2B_OK;
    if (bytes < STAGE_MIN_BYTES || !host_is_pageable(h_src)) {
        H2B_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, consumer));
        return H2B_OK;
    }
    H2B_TRY(stager_init(ctx));
    Stager& s = *ctx.stager;
    if (order_after_consumer) H2B_CUDA(cudaEventRecord(s.ready, consumer));
    const size_t pieces = (bytes + STAGE_PIECE - 1) / STAGE_PIECE;
    const int T = (int)(pieces < (size_t)s.nthreads ? pieces : (size_t)s.nthreads);
    std::vector<cudaError_t> errs(T, cudaSuccess);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) {
        th.emplace_back([&, t] {
            cudaError_t e = cudaSetDevice(ctx.device);
            if (e == cudaSuccess && order_after_consumer) e = cudaStreamWaitEvent(s.stream[t], s.ready, 0);
            size_t k = 0;
            for (size_t p = t; p < pieces && e == cudaSuccess; p += T, ++k) {
                const int sl = (int)(k % STAGE_SLOTS);
                const size_t off = p * STAGE_PIECE, len = (bytes - off < STAGE_PIECE) ? bytes - off : STAGE_PIECE;
                // the slot's previous DMA -- of this call or of an earlier one, which may return before its DMAs finish --
                // has drained (a never-recorded event synchronises immediately)
                e = cudaEventSynchronize(s.slot_free[t][sl]);
                if (e != cudaSuccess) break;
                memcpy(s.slot[t][sl], (const char*)h_src + off, len);
                e = cudaMemcpyAsync((char*)d_dst + off, s.slot[t][sl], len, cudaMemcpyHostToDevice, s.stream[t]);
                if (e == cudaSuccess) e = cudaEventRecord(s.slot_free[t][sl], s.stream[t]);
            }
            if (e == cudaSuccess) e = cudaEventRecord(s.done[t], s.stream[t]);
            errs[t] = e;
        });
    }
    for (auto& x : th) x.join();
    for (int t = 0; t < T; ++t) H2B_CUDA(errs[t]);
    for (int t = 0; t < T; ++t) H2B_CUDA(cudaStreamWaitEvent(consumer, s.done[t], 0));
    return H2B_OK;
}

// Synchronous: returns when h_dst holds the bytes that `producer` had written to d_src.
int host_download(DeviceCtx& ctx, void* h_dst, const void* d_src, size_t bytes, cudaStream_t producer) {
    if (bytes == 0) return H2B_OK;
    if (bytes < STAGE_MIN_BYTES || !host_is_pageable(h_dst)) {
        H2B_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, producer));
        H2B_CUDA(cudaStreamSynchronize(producer));
        return H2B_OK;
    }
    H2B_TRY(stager_init(ctx));
    Stager& s = *ctx.stager;
    H2B_CUDA(cudaEventRecord(s.ready, producer));
    const size_t pieces = (bytes + STAGE_PIECE - 1) / STAGE_PIECE;
    const int T = (int)(pieces < (size_t)s.nthreads ? pieces : (size_t)s.nthreads);
    std::vector<cudaError_t> errs(T, cudaSuccess);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) {
        th.emplace_back([&, t] {
            cudaError_t e = cudaSetDevice(ctx.device);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s.stream[t], s.ready, 0);
            // software pipeline over this worker's pieces: DMA of piece k + 1 is in flight while piece k is copied out
            auto issue = [&](size_t k) -> cudaError_t {
                const size_t p = t + k * (size_t)T;
                const int sl = (int)(k % STAGE_SLOTS);
                const size_t off = p * STAGE_PIECE, len = (bytes - off < STAGE_PIECE) ? bytes - off : STAGE_PIECE;
                cudaError_t r = cudaMemcpyAsync(s.slot[t][sl], (const char*)d_src + off, len, cudaMemcpyDeviceToHost, s.stream[t]);
                if (r == cudaSuccess) r = cudaEventRecord(s.slot_free[t][sl], s.stream[t]);
                return r;
            };
            const size_t mine = (pieces > (size_t)t) ? (pieces - t + T - 1) / T : 0;
            if (e == cudaSuccess && mine > 0) e = issue(0);
            for (size_t k = 0; k < mine && e == cudaSuccess; ++k) {
                if (k + 1 < mine) e = issue(k + 1);
                if (e != cudaSuccess) break;
                const size_t p = t + k * (size_t)T;
                const int sl = (int)(k % STAGE_SLOTS);
                const size_t off = p * STAGE_PIECE, len = (bytes - off < STAGE_PIECE) ? bytes - off : STAGE_PIECE;
                e = cudaEventSynchronize(s.slot_free[t][sl]);
                if (e == cudaSuccess) memcpy((char*)h_dst + off, s.slot[t][sl], len);
            }
            errs[t] = e;
        });
    }
    for (auto& x : th) x.join();
    for (int t = 0; t < T; ++t) H2B_CUDA(errs[t]);
    return H2B_OK;
}

// ---- strided host arrays (the row / column blocks of the multi-device NTT) ------------------------------------------------------
// `height` rows of `width` bytes, `h_pitch` bytes apart in host memory, packed on the device.  Pinned host memory: one 2-D DMA.
// Pageable: the staging threads gather whole rows into their pinned slots (a piece = as many rows as fit one slot) and DMA them.
int host_upload_2d(DeviceCtx& ctx, void* d_dst, const void* h_src, size_t h_pitch, size_t width, size_t height, cudaStream_t consumer) {
    if (width == 0 || height == 0) return H2B_OK;
    if (h_pitch == width) return host_upload(ctx, d_dst, h_src, width * height, consumer);
    if (!host_is_pageable(h_src) || width > STAGE_PIECE || width * height < STAGE_MIN_BYTES) {
        H2B_CUDA(cudaMemcpy2DAsync(d_dst, width, h_src, h_pitch, width, height, cudaMemcpyHostToDevice, consumer));
        return H2B_OK;
    }
    H2B_TRY(stager_init(ctx));
    Stager& s = *ctx.stager;
    H2B_CUDA(cudaEventRecord(s.ready, consumer));
    const size_t rows_per_piece = STAGE_PIECE / width;
    const size_t pieces = (height + rows_per_piece - 1) / rows_per_piece;
    const int T = (int)(pieces < (size_t)s.nthreads ? pieces : (size_t)s.nthreads);
    std::vector<cudaError_t> errs(T, cudaSuccess);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) {
        th.emplace_back([&, t] {
            cudaError_t e = cudaSetDevice(ctx.device);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s.stream[t], s.ready, 0);
            size_t k = 0;
            for (size_t p = t; p < pieces && e == cudaSuccess; p += T, ++k) {
                const int sl = (int)(k % STAGE_SLOTS);
                const size_t r0 = p * rows_per_piece, r1 = r0 + rows_per_piece < height ? r0 + rows_per_piece : height;
                e = cudaEventSynchronize(s.slot_free[t][sl]);
                if (e != cudaSuccess) break;
                for (size_t r = r0; r < r1; ++r) memcpy((char*)s.slot[t][sl] + (r - r0) * width, (const char*)h_src + r * h_pitch, width);
                e = cudaMemcpyAsync((char*)d_dst + r0 * width, s.slot[t][sl], (r1 - r0) * width, cudaMemcpyHostToDevice, s.stream[t]);
                if (e == cudaSuccess) e = cudaEventRecord(s.slot_free[t][sl], s.stream[t]);
            }
            if (e == cudaSuccess) e = cudaEventRecord(s.done[t], s.stream[t]);
            errs[t] = e;
        });
    }
    for (auto& x : th) x.join();
    for (int t = 0; t < T; ++t) H2B_CUDA(errs[t]);
    for (int t = 0; t < T; ++t) H2B_CUDA(cudaStreamWaitEvent(consumer, s.done[t], 0));
    return H2B_OK;
}

// Synchronous, like host_download.
int host_download_2d(DeviceCtx& ctx, void* h_dst, size_t h_pitch, const void* d_src, size_t width, size_t height, cudaStream_t producer) {
    if (width == 0 || height == 0) return H2B_OK;
    if (h_pitch == width) return host_download(ctx, h_dst, d_src, width * height, producer);
    if (!host_is_pageable(h_dst) || width > STAGE_PIECE || width * height < STAGE_MIN_BYTES) {
        H2B_CUDA(cudaMemcpy2DAsync(h_dst, h_pitch, d_src, width, width, height, cudaMemcpyDeviceToHost, producer));
        H2B_CUDA(cudaStreamSynchronize(producer));
        return H2B_OK;
    }
    H2B_TRY(stager_init(ctx));
    Stager& s = *ctx.stager;
    H2B_CUDA(cudaEventRecord(s.ready, producer));
    const size_t rows_per_piece = STAGE_PIECE / width;
    const size_t pieces = (height + rows_per_piece - 1) / rows_per_piece;
    const int T = (int)(pieces < (size_t)s.nthreads ? pieces : (size_t)s.nthreads);
    std::vector<cudaError_t> errs(T, cudaSuccess);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) {
        th.emplace_back([&, t] {
            cudaError_t e = cudaSetDevice(ctx.device);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s.stream[t], s.ready, 0);
            auto rows_of = [&](size_t k, size_t& r0, size_t& r1) {
                const size_t p = t + k * (size_t)T;
                r0 = p * rows_per_piece;
                r1 = r0 + rows_per_piece < height ? r0 + rows_per_piece : height;
            };
            auto issue = [&](size_t k) -> cudaError_t {
                size_t r0, r1;
                rows_of(k, r0, r1);
                const int sl = (int)(k % STAGE_SLOTS);
                cudaError_t r = cudaMemcpyAsync(s.slot[t][sl], (const char*)d_src + r0 * width, (r1 - r0) * width, cudaMemcpyDeviceToHost, s.stream[t]);
                if (r == cudaSuccess) r = cudaEventRecord(s.slot_free[t][sl], s.stream[t]);
                return r;
            };
            const size_t mine = (pieces > (size_t)t) ? (pieces - t + T - 1) / T : 0;
            if (e == cudaSuccess && mine > 0) e = issue(0);
            for (size_t k = 0; k < mine && e == cudaSuccess; ++k) {
                if (k + 1 < mine) e = issue(k + 1);
                if (e != cudaSuccess) break;
                size_t r0, r1;
                rows_of(k, r0, r1);
                const int sl = (int)(k % STAGE_SLOTS);
                e = cudaEventSynchronize(s.slot_free[t][sl]);
                if (e == cudaSuccess)
                    for (size_t r = r0; r < r1; ++r) memcpy((char*)h_dst + r * h_pitch, (const char*)s.slot[t][sl] + (r - r0) * width, width);
            }
            errs[t] = e;
        });
    }
    for (auto& x : th) x.join();
    for (int t = 0; t < T; ++t) H2B_CUDA(errs[t]);
    return H2B_OK;
}

}  // namespace h2b
