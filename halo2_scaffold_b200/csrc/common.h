// Shared host-side plumbing of libh2b200: per-device context, error reporting, scratch buffers.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/h2b200.h"   // H2B_OK / H2B_ERR_* codes
#include "field.cuh"

namespace h2b {

void set_error(const char* fmt, ...);
const char* get_error();

#define H2B_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            h2b::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? H2B_ERR_OOM : H2B_ERR_CUDA;     \
        }                                                                                        \
    } while (0)

#define H2B_TRY(expr)             \
    do {                          \
        int _r = (expr);          \
        if (_r != 0) return _r;   \
    } while (0)

// A grow-only device buffer (scratch space reused across calls on one device).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return H2B_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            return H2B_ERR_OOM;
        }
        cap = want;
        return H2B_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers).
enum ProfTag { PROF_BEGIN = 0, PROF_MSM_DECOMPOSE = 1, PROF_MSM_SCAN = 2, PROF_MSM_SCATTER = 3, PROF_MSM_PLAN = 4, PROF_MSM_ACCUMULATE = 5,
               PROF_MSM_COMBINE = 6, PROF_MSM_REDUCE = 7, PROF_MSM_FINAL = 8, PROF_MSM_PRECOMPUTE = 9, PROF_MSM_PAIR = 10, PROF_NTT_TWIDDLE = 15, PROF_NTT_PASS0 = 16 };
struct Profiler {
    bool enabled = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> tag;
    size_t used = 0;
    void mark(int t, cudaStream_t stream) {
        if (!enabled || used >= ev.size()) return;
        cudaEventRecord(ev[used], stream);
        tag[used++] = t;
    }
};

struct Stager;        // stage.cu
struct NttTwiddles;   // ntt.cu
struct MsmScratch;    // msm.cu
struct BaseSet;       // api.cu
struct EvalScratch;   // evaluate.cu

struct DeviceCtx {
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;     // stream used by the host-pointer entry points
    cudaStream_t copy_stream = nullptr;   // chunked scalar uploads of the host-pointer MSM (overlap with compute)
    std::vector<cudaEvent_t> copy_events;
    std::mutex mu;                     // serialises calls that share this context's scratch
    DevBuf ntt_work;                   // ping-pong buffer of the multi-pass NTT
    bool ntt_attr_set = false;
    bool msm_attr_set = false;
    DevBuf ntt_io;                     // staging for the host-pointer NTT entry point
    DevBuf ntt_fs[2];                  // the two other matrices of the multi-device four-step NTT (api.cu: ntt_multi_device)
    DevBuf col_ext;                    // extended form of h2b_column_pipeline when the caller only wants it on the host
    DevBuf scale_table;                // factor table of h2b_fr_scale_dev with more than 8 factors
    DevBuf msm_scalars;                // staging for host-pointer MSM scalars
    DevBuf msm_out;                    // 96-byte result
    DevBuf msm_bases;                  // points of an uncached host-pointer MSM (small / odd-length calls: uploaded per call)
    DevBuf scan_scratch;               // batch inversion / prefix product scratch (scan.cu)
    DevBuf srs_status, srs_io;         // first-invalid-point flag and encoded-bytes staging of the SRS reader (srs.cu)
    DevBuf lookup_scratch;             // sorted keys, flags and ranks of permute_expression_pair (lookup.cu)
    EvalScratch* eval = nullptr;       // compiled-program ring of the quotient evaluation (evaluate.cu)
    std::vector<NttTwiddles*> twiddles;  // small LRU cache keyed by (omega, log_n)
    MsmScratch* msm = nullptr;
    Profiler prof;
    Stager* stager = nullptr;          // pinned slots + worker streams for pageable host buffers
    void* pinned = nullptr;            // small pinned bounce buffer
    size_t pinned_cap = 0;
};

// ---- stage.cu ----
int host_upload(DeviceCtx& ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t consumer, bool order_after_consumer = true);
int host_download(DeviceCtx& ctx, void* h_dst, const void* d_src, size_t bytes, cudaStream_t producer);
// `height` rows of `width` bytes: host rows are `h_pitch` bytes apart, the device side is packed
int host_upload_2d(DeviceCtx& ctx, void* d_dst, const void* h_src, size_t h_pitch, size_t width, size_t height, cudaStream_t consumer);
int host_download_2d(DeviceCtx& ctx, void* h_dst, size_t h_pitch, const void* d_src, size_t width, size_t height, cudaStream_t producer);
void stager_release(DeviceCtx& ctx);
bool host_is_pageable(const void* p);
// ---- scan.cu ----
int fr_batch_invert_run(DeviceCtx& ctx, void* d_a, size_t n, cudaStream_t stream);
int fr_prefix_product_run(DeviceCtx& ctx, const void* d_in, void* d_out, size_t n, cudaStream_t stream);
int fr_eval_polynomial_run(DeviceCtx& ctx, const void* d_coeffs, size_t n, const uint64_t x[4], void* d_out, cudaStream_t stream);
int fr_kate_division_run(DeviceCtx& ctx, const void* d_a, size_t n, const uint64_t b[4], void* d_q, cudaStream_t stream);
int fr_lincomb_run(DeviceCtx& ctx, const void* const* d_cols, const uint64_t* coeffs, uint32_t m, size_t n, void* d_out, cudaStream_t stream);
int permutation_product_run(DeviceCtx& ctx, const void* const* d_values, const void* const* d_sigma, uint32_t m, size_t n, const uint64_t* beta,
                            const uint64_t* gamma, const uint64_t* delta, const uint64_t* deltaomega, const uint64_t* omega, const uint64_t* last_z,
                            void* d_z, cudaStream_t stream);
int lookup_product_run(DeviceCtx& ctx, const void* d_compressed_input, const void* d_compressed_table, const void* d_permuted_input,
                       const void* d_permuted_table, size_t n, const uint64_t* beta, const uint64_t* gamma, void* d_z, cudaStream_t stream);
// ---- lookup.cu ----
int lookup_permute_run(DeviceCtx& ctx, const void* d_input, const void* d_table, uint32_t usable_rows, void* d_permuted_input, void* d_permuted_table,
                       void* d_status, cudaStream_t stream);
// ---- srs.cu ----
int g1_decode_run(DeviceCtx& ctx, const void* d_bytes, size_t n, int format, void* d_out, uint64_t* first_invalid, cudaStream_t stream);
int g1_encode_run(DeviceCtx& ctx, const void* d_affine, size_t n, void* d_out_bytes, cudaStream_t stream);
// ---- evaluate.cu ----
int evaluate_graph_run(DeviceCtx& ctx, const h2b_graph* g, const h2b_eval_columns* cols, void* d_values, uint32_t size, int32_t rot_scale,
                       const h2b_eval_shard* shard, cudaStream_t stream);
int evaluate_h_lookup_run(DeviceCtx& ctx, const h2b_graph* g, const h2b_eval_columns* cols, void* d_values, uint32_t size, int32_t rot_scale,
                          const void* d_product, const void* d_permuted_input, const void* d_permuted_table, const void* d_l0, const void* d_l_last,
                          const void* d_l_active_row, const h2b_eval_shard* shard, cudaStream_t stream);
int evaluate_h_permutation_run(DeviceCtx& ctx, void* d_values, uint32_t size, int32_t rot_scale, const void* const* d_product_cosets, uint32_t n_sets,
                               const void* const* d_columns, const void* const* d_perm_cosets, uint32_t n_columns, uint32_t chunk_len, int32_t last_rotation,
                               const void* d_l0, const void* d_l_last, const void* d_l_active_row, const uint64_t* beta, const uint64_t* gamma,
                               const uint64_t* y, const uint64_t* delta, const uint64_t* zeta, const uint64_t* extended_omega, const h2b_eval_shard* shard,
                               cudaStream_t stream);
void evaluate_graph_last_info(uint32_t* slots, uint32_t* micro_ops);
void evaluate_release(DeviceCtx& ctx);
// ---- ntt.cu ----
int ntt_run(DeviceCtx& ctx, void* d_a, const uint64_t omega[4], uint32_t log_n, cudaStream_t stream);
int ntt_run_batch(DeviceCtx& ctx, void* const* d_polys, size_t count, const uint64_t omega[4], uint32_t log_n, cudaStream_t stream);
uint32_t ntt_batch_max();
int ntt_run_strided(DeviceCtx& ctx, void* d_base, size_t count, size_t stride_elems, const uint64_t omega[4], uint32_t log_n, cudaStream_t stream);
int fr_transpose_run(DeviceCtx& ctx, const void* d_in, void* d_out, uint32_t rows, uint32_t cols, cudaStream_t stream);
int ntt_fourstep_twiddle_run(DeviceCtx& ctx, void* d_y, uint32_t rows, uint32_t log_len, uint32_t row0, const uint64_t omega[4], cudaStream_t stream);
int ntt_root_powers_run(DeviceCtx& ctx, const uint64_t omega[4], uint32_t e0, uint32_t e1, void* d_out, cudaStream_t stream);
int ntt_scale_run(DeviceCtx& ctx, void* d_a, size_t n, const uint64_t* factors /*host, count x 4*/, int count, cudaStream_t stream);
int ntt_scale_batch_run(DeviceCtx& ctx, void* const* d_cols, size_t ncols, size_t n, const uint64_t* factors, int count, cudaStream_t stream);
void ntt_release(DeviceCtx& ctx);
// ---- msm.cu ----
// The points of an MSM: table j (j < n_tables) holds 2^(c0*j) * P_i at rows [0, stride); the call uses rows
// [row0, row0 + n) of every table.  n_tables == 1 is the plain, table-less mode (c0 ignored).
struct MsmBases {
    const void* tables = nullptr;
    uint32_t n_tables = 1;
    uint32_t c0 = 0;
    size_t stride = 0;
    size_t row0 = 0;
};
int msm_run(DeviceCtx& ctx, const void* d_scalars, const MsmBases& bases, size_t n, void* d_out, bool with_xyzz, cudaStream_t stream);
int msm_run_host(DeviceCtx& ctx, const void* h_scalars, void* d_staging, const MsmBases& bases, size_t n, void* h_out_block);
int msm_run_batch(DeviceCtx& ctx, const void* const* d_cols, const size_t* lens, uint32_t count, const MsmBases& bases, void* d_out_blocks, cudaStream_t stream);
int msm_run_host_batch(DeviceCtx& ctx, const void* const* h_cols, const size_t* lens, uint32_t count, void* d_staging, const MsmBases& bases, void* h_out_blocks);
uint32_t msm_batch_max();
int msm_precompute_run(DeviceCtx& ctx, const void* d_src, void* d_dst, size_t n, uint32_t c0, cudaStream_t stream);
uint32_t msm_pick_table_spacing(size_t n, uint32_t max_tables);
uint32_t msm_tables_for(uint32_t c0);
int msm_sum_partials_run(DeviceCtx& ctx, const void* d_blocks, uint32_t count, void* d_out_jac, cudaStream_t stream);
void msm_release(DeviceCtx& ctx);
int msm_set_window(int c);   // 0 = automatic
// ---- testgen.cu ----
int gen_points_run(DeviceCtx& ctx, uint64_t seed, size_t n, void* d_out_affine, cudaStream_t stream);
int msm_checksum_run(DeviceCtx& ctx, const void* d_scalars, uint64_t seed, uint64_t first, size_t n, void* d_out, cudaStream_t stream);
int gen_scalars_run(DeviceCtx& ctx, uint64_t seed, size_t n, int kind, void* d_out, cudaStream_t stream);
int field_selftest_run(DeviceCtx& ctx, int field, int op, const void* d_a, const void* d_b, size_t n, void* d_out, cudaStream_t stream);
int ec_selftest_run(DeviceCtx& ctx, int op, const void* d_p, const void* d_q, size_t n, void* d_out, cudaStream_t stream);
int imad_bench_run(DeviceCtx& ctx, int kind, int iters, int blocks, int threads, float* ms_out, double* ops_out, cudaStream_t stream);

}  // namespace h2b
