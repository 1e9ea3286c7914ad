// tools.cu -> libtwkb_tools.so: measurement tooling of bench.py / tests. NOT part of the product path
// (libtwkb.so never links or loads it; nothing here computes LD):
//   * twkb_tools_synth      synthetic genotype matrix generated ON THE DEVICE, in the row layout
//                           twkb_load_matrix_device takes (SURVEY.md 8d "synthetic inputs": LD blocks, skewed
//                           allele-frequency spectrum, optional rare-variant share and missing genotypes). numpy
//                           needs ~80 s per GB; BASELINE configs[3] is 12.5 GB per GPU.
//   * twkb_tools_popc_rate  INT-pipe POPC issue rate (the roofline denominator of count_popc_kernel).
//   * twkb_tools_fp4_gemm   block-scaled e2m1 GEMM through cuBLASLt: a measured tensor-pipe denominator for the
//                           e2m1 count kernel (MEASURED_PEAKS.json only carries bf16).
#include <cublasLt.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/twkb.h"

namespace {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finalizer
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
__host__ __device__ __forceinline__ uint64_t key3(uint64_t seed, uint64_t v, uint64_t k) {
    return mix64(seed * 0x9e3779b97f4a7c15ull + v * 0xd1342543de82ef95ull + k * 0x2545f4914f6cdd1dull + 0x632be59bd9b4e019ull);
}
__device__ __forceinline__ float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

struct SynthPrm {
    uint64_t seed;
    uint32_t n_samples, n_variants, first_variant;
    float p_copy, redraw, rare_fraction, missing_rate;
    uint32_t pos_step;
};

// LD-block structure: variant v starts a block with probability 1 - p_copy (and at every multiple of 64, which
// bounds the backward scan); a block member copies its founder's haplotypes and re-draws each one with
// probability 1 - (1 - redraw)^depth.
__device__ __forceinline__ bool starts_block(const SynthPrm& p, uint32_t v) {
    return v == 0 || (v & 63u) == 0u || u01(key3(p.seed, v, 0)) >= p.p_copy;
}

__global__ void synth_rows_kernel(SynthPrm p, uint64_t* __restrict__ data, uint64_t* __restrict__ mask, size_t stride) {
    const uint32_t vl = blockIdx.x;                       // row inside this call
    const uint32_t v = p.first_variant + vl;              // global variant index (what the random stream is keyed on)
    const uint32_t H = 2u * p.n_samples;
    __shared__ uint32_t s_founder;
    __shared__ float s_af, s_pr;
    if (threadIdx.x == 0) {
        uint32_t f = v;
        while (!starts_block(p, f)) --f;
        float af = 0.5f * powf(u01(key3(p.seed, f, 1)), 3.0f);
        if (p.rare_fraction > 0.0f && u01(key3(p.seed, f, 2)) < p.rare_fraction) af = 0.01f * u01(key3(p.seed, f, 3));
        af = fmaxf(af, 1.0f / (float)H);
        s_founder = f;
        s_af = af;
        s_pr = 1.0f - powf(1.0f - p.redraw, (float)(v - f));
    }
    __syncthreads();
    const uint32_t f = s_founder;
    const uint32_t af_u = (uint32_t)fminf(s_af * 4294967296.0f, 4294967295.0f);
    const uint32_t pr_u = (uint32_t)fminf(s_pr * 4294967296.0f, 4294967295.0f);
    const uint32_t miss_u = (uint32_t)fminf(p.missing_rate * 4294967296.0f, 4294967295.0f);
    for (uint32_t w = threadIdx.x; w < (uint32_t)stride; w += blockDim.x) {
        uint64_t bits = 0, mbits = 0;
        const uint32_t h0 = w * 64u;
        if (h0 < H) {
            const uint32_t n = min(64u, H - h0);
            for (uint32_t b = 0; b < n; ++b) {
                const uint32_t h = h0 + b;
                const uint64_t rf = key3(p.seed, f, 16ull + h);
                bool alt = (uint32_t)rf < af_u;
                if (v != f) {
                    const uint64_t rm = key3(p.seed, v, 16ull + h);
                    if ((uint32_t)rm < pr_u) alt = (uint32_t)(rm >> 32) < af_u;
                }
                bits |= (uint64_t)alt << b;
            }
            if (mask) {  // a sample is missing as a whole (both alleles), lib/core.cpp:379-380
                for (uint32_t b = 0; b < n; b += 2) {
                    const uint32_t s = (h0 + b) >> 1;
                    if ((uint32_t)key3(p.seed, v, (1ull << 40) + s) < miss_u) mbits |= 3ull << b;
                }
                bits &= ~mbits;  // data bits are 0 where missing (lib/core.cpp:377-380)
            }
        }
        data[(size_t)vl * stride + w] = bits;
        if (mask) mask[(size_t)vl * stride + w] = mbits;
    }
}

// Allele counts + the reference's constraint 1 <= ac (a monomorphic site trips assert(list != nullptr),
// include/core.h:547) and ac <= (non-missing haplotypes) - 1, then the metadata record.
__global__ void synth_meta_kernel(SynthPrm p, uint64_t* __restrict__ data, const uint64_t* __restrict__ mask, size_t stride,
                                  twkb_variant* __restrict__ meta) {
    const uint32_t vl = blockIdx.x, v = p.first_variant + vl;
    uint64_t* row = data + (size_t)vl * stride;
    const uint64_t* mrow = mask ? mask + (size_t)vl * stride : nullptr;
    uint32_t ac = 0, an = 0;
    for (uint32_t w = threadIdx.x; w < (uint32_t)stride; w += blockDim.x) {
        ac += __popcll(row[w]);
        if (mrow) an += __popcll(mrow[w]);
    }
    __shared__ uint32_t s_ac[32], s_an[32];
    for (int d = 16; d > 0; d >>= 1) { ac += __shfl_down_sync(0xffffffffu, ac, d); an += __shfl_down_sync(0xffffffffu, an, d); }
    if ((threadIdx.x & 31) == 0) { s_ac[threadIdx.x >> 5] = ac; s_an[threadIdx.x >> 5] = an; }
    __syncthreads();
    if (threadIdx.x == 0) {
        ac = an = 0;
        for (uint32_t i = 0; i < (blockDim.x + 31) / 32; ++i) { ac += s_ac[i]; an += s_an[i]; }
        const uint32_t H = 2u * p.n_samples;
        if (ac == 0) {  // one carrier at a pseudo-random haplotype (the next non-missing one)
            const uint32_t h0 = (uint32_t)(key3(p.seed, v, 7) % H);
            for (uint32_t k = 0; k < H; ++k) {
                const uint32_t h = (h0 + k) % H;
                if (!mrow || !((mrow[h >> 6] >> (h & 63)) & 1ull)) { row[h >> 6] |= 1ull << (h & 63); ac = 1; break; }
            }
        } else if (ac == H - an && ac > 1) {  // clear the first alt haplotype
            for (uint32_t h = 0; h < H; ++h)
                if ((row[h >> 6] >> (h & 63)) & 1ull) { row[h >> 6] &= ~(1ull << (h & 63)); --ac; break; }
        }
        twkb_variant m;
        memset(&m, 0, sizeof(m));
        m.rid = 0;
        m.pos = v * p.pos_step;
        m.ac = ac;
        m.an = an;
        m.hwe = 1.0;
        m.gt_missing = an ? 1 : 0;
        m.gt_phase = 1;
        meta[vl] = m;
    }
}

// ---- POPC issue rate: 64 DISTINCT popc(a_i & b_j) per iteration per thread (8 x 8 register tile, the shape of
// count_popc_kernel's inner loop), operands in registers, no memory traffic
__global__ void popc_rate_kernel(uint32_t iters, uint32_t seed, uint32_t* out) {
    uint32_t a[8], b[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed * (2654435761u + 40503u * i) + threadIdx.x + blockIdx.x * 977u;
        b[i] = seed * (2246822519u + 30011u * i) ^ (threadIdx.x * 31u + blockIdx.x);
        acc[i] = 0;
    }
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += __popc(a[i] & b[j]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += acc[i]; b[i] ^= acc[(i + 3) & 7]; }  // keeps the compiler from hoisting the popcounts
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= acc[i];
    if (x == 0x12345678u) out[0] = x;
}

std::string g_err;

}  // namespace

extern "C" {

const char* twkb_tools_last_error(void) { return g_err.c_str(); }

// Fills rows [first_variant, first_variant + n_rows) of the synthetic matrix of `n_variants` x `n_samples`
// (the random stream is keyed on the global variant index: any slice of the same (seed, shape) is the same data).
// d_data / d_mask (nullable, required when missing_rate > 0): device pointers, n_rows x stride u64 words;
// h_meta: host array of n_rows entries. Returns 0 or a negative cudaError.
int twkb_tools_synth(uint64_t seed, uint32_t n_samples, uint32_t n_variants, uint32_t first_variant, uint32_t n_rows,
                     double p_copy, double redraw, double rare_fraction, double missing_rate, uint32_t pos_step, uint64_t* d_data,
                     uint64_t* d_mask, size_t stride, twkb_variant* h_meta) {
    if (!d_data || !h_meta || n_rows == 0 || first_variant + (uint64_t)n_rows > n_variants || stride * 64 < 2ull * n_samples ||
        (missing_rate > 0 && !d_mask)) {
        g_err = "twkb_tools_synth: bad arguments";
        return -1;
    }
    SynthPrm p{seed, n_samples, n_variants, first_variant, (float)p_copy, (float)redraw, (float)rare_fraction, (float)missing_rate, pos_step};
    synth_rows_kernel<<<n_rows, 256>>>(p, d_data, d_mask, stride);
    twkb_variant* d_meta = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_meta, (size_t)n_rows * sizeof(twkb_variant));
    if (e == cudaSuccess) {
        synth_meta_kernel<<<n_rows, 256>>>(p, d_data, d_mask, stride, d_meta);
        e = cudaMemcpy(h_meta, d_meta, (size_t)n_rows * sizeof(twkb_variant), cudaMemcpyDeviceToHost);
        cudaFree(d_meta);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { g_err = std::string("twkb_tools_synth: ") + cudaGetErrorString(e); return -(int)e; }
    return 0;
}

// POPC instructions per clock per SM and per second on the current device (AND + POPC + ADD per word operation,
// every SM saturated with 2048 threads). Returns 0 on success.
int twkb_tools_popc_rate(double* popc_per_s, double* popc_per_clk_per_sm, double* sm_mhz_effective) {
    int dev = 0, n_sm = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    uint32_t* d_out = nullptr;
    if (cudaMalloc((void**)&d_out, 4) != cudaSuccess) { g_err = "cudaMalloc"; return -1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const uint32_t iters = 20000;
    const int blocks = n_sm * 8, threads = 256;
    popc_rate_kernel<<<blocks, threads>>>(200, 1, d_out);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        popc_rate_kernel<<<blocks, threads>>>(iters, 2 + rep, d_out);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { g_err = "popc_rate_kernel failed"; cudaFree(d_out); return -1; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)blocks * threads * (double)iters * 64.0;
    if (popc_per_s) *popc_per_s = ops / (best * 1e-3);
    if (popc_per_clk_per_sm) *popc_per_clk_per_sm = ops / (best * 1e-3) / ((double)khz * 1e3) / n_sm;
    if (sm_mhz_effective) *sm_mhz_effective = khz / 1e3;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    return 0;
}

// Block-scaled e2m1 x e2m1 -> bf16 GEMM (NVFP4: 16-element blocks, UE4M3 scales) of size n^3 through cuBLASLt.
// burst = best of 10 launches, sustained = back to back for `sustain_s` seconds. TFLOP/s counted as 2 n^3.
int twkb_tools_fp4_gemm(uint32_t n, double sustain_s, double* tflops_burst, double* tflops_sustained) {
    if (n == 0 || n % 128) { g_err = "n must be a multiple of 128"; return -1; }
    cublasLtHandle_t lt = nullptr;
    if (cublasLtCreate(&lt) != CUBLAS_STATUS_SUCCESS) { g_err = "cublasLtCreate failed"; return -1; }
    void *A = nullptr, *B = nullptr, *C = nullptr, *sA = nullptr, *sB = nullptr, *ws = nullptr;
    const size_t elems = (size_t)n * n, ws_bytes = 64u << 20;
    auto cleanup = [&] {
        cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(sA); cudaFree(sB); cudaFree(ws);
        if (lt) cublasLtDestroy(lt);
    };
    if (cudaMalloc(&A, elems / 2) || cudaMalloc(&B, elems / 2) || cudaMalloc(&C, elems * 2) || cudaMalloc(&sA, elems / 16 + 4096) ||
        cudaMalloc(&sB, elems / 16 + 4096) || cudaMalloc(&ws, ws_bytes)) {
        g_err = "cudaMalloc failed";
        cleanup();
        return -1;
    }
    cudaMemset(A, 0x22, elems / 2);   // every element e2m1 1.0
    cudaMemset(B, 0x22, elems / 2);
    cudaMemset(sA, 0x38, elems / 16 + 4096);  // UE4M3 1.0
    cudaMemset(sB, 0x38, elems / 16 + 4096);
    cublasLtMatmulDesc_t op = nullptr;
    cublasLtMatrixLayout_t la = nullptr, lb = nullptr, lc = nullptr;
    cublasLtMatmulPreference_t pref = nullptr;
    int rc = -1;
    do {
        if (cublasLtMatmulDescCreate(&op, CUBLAS_COMPUTE_32F, CUDA_R_32F) != CUBLAS_STATUS_SUCCESS) { g_err = "MatmulDescCreate"; break; }
        cublasOperation_t T = CUBLAS_OP_T, N = CUBLAS_OP_N;
        cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &T, sizeof(T));
        cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &N, sizeof(N));
        cublasLtMatmulMatrixScale_t mode = CUBLASLT_MATMUL_MATRIX_SCALE_VEC16_UE4M3;
        if (cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_A_SCALE_MODE, &mode, sizeof(mode)) != CUBLAS_STATUS_SUCCESS ||
            cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_B_SCALE_MODE, &mode, sizeof(mode)) != CUBLAS_STATUS_SUCCESS) {
            g_err = "block-scale mode not supported by this cuBLASLt";
            break;
        }
        cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_A_SCALE_POINTER, &sA, sizeof(sA));
        cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_B_SCALE_POINTER, &sB, sizeof(sB));
        if (cublasLtMatrixLayoutCreate(&la, CUDA_R_4F_E2M1, n, n, n) != CUBLAS_STATUS_SUCCESS ||
            cublasLtMatrixLayoutCreate(&lb, CUDA_R_4F_E2M1, n, n, n) != CUBLAS_STATUS_SUCCESS ||
            cublasLtMatrixLayoutCreate(&lc, CUDA_R_16BF, n, n, n) != CUBLAS_STATUS_SUCCESS) {
            g_err = "MatrixLayoutCreate (e2m1)";
            break;
        }
        cublasLtMatmulPreferenceCreate(&pref);
        cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes));
        cublasLtMatmulHeuristicResult_t heur{};
        int found = 0;
        if (cublasLtMatmulAlgoGetHeuristic(lt, op, la, lb, lc, lc, pref, 1, &heur, &found) != CUBLAS_STATUS_SUCCESS || found == 0) {
            g_err = "cuBLASLt has no block-scaled e2m1 algorithm for this shape";
            break;
        }
        const float alpha = 1.0f, beta = 0.0f;
        auto launch = [&]() { return cublasLtMatmul(lt, op, &alpha, A, la, B, lb, &beta, C, lc, C, lc, &heur.algo, ws, ws_bytes, 0); };
        if (launch() != CUBLAS_STATUS_SUCCESS || cudaDeviceSynchronize() != cudaSuccess) { g_err = "cublasLtMatmul (e2m1) failed"; break; }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const double flop = 2.0 * (double)n * n * n;
        float best = 1e30f;
        for (int i = 0; i < 10; ++i) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        if (tflops_burst) *tflops_burst = flop / (best * 1e-3) / 1e12;
        const int reps = (int)(sustain_s / (best * 1e-3)) + 1;
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (tflops_sustained) *tflops_sustained = flop * reps / (ms * 1e-3) / 1e12;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        rc = 0;
    } while (0);
    if (pref) cublasLtMatmulPreferenceDestroy(pref);
    if (la) cublasLtMatrixLayoutDestroy(la);
    if (lb) cublasLtMatrixLayoutDestroy(lb);
    if (lc) cublasLtMatrixLayoutDestroy(lc);
    if (op) cublasLtMatmulDescDestroy(op);
    cleanup();
    return rc;
}

}  // extern "C"
