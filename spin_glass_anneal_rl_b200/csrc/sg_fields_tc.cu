// sg_fields_tc.cu -- K2-TC: local-field initialisation F = S J^T + h on the int8 tensor cores.
//
// Restates R x IsingModel.get_local_field / compute_energy (reference core/ising_model.py:149-185)
// and BatchProcessor.process_batch_energies / VectorizedOperations.vectorized_local_fields
// (optimization/high_performance_computing.py:98-165, 357-372) as ONE exact GEMM.
//
// Exactness.  The couplings are converted once per model to fixed point, q = rint(J * 2^s) with
// |q| < 2^31 (s = 0 when every coupling is an integer), and q is split into four balanced base-256
// digits a_d in [-128, 127], q = sum_d a_d 256^d.  Spins are +-1 int8.  tcgen05.mma.kind::i8
// accumulates sum_i a_d(i, j) s_ri exactly in int32 (|sum| <= 4096 * 128), one accumulator per
// digit; the epilogue recombines the four int32 sums in int64 and scales by 2^-s in double.  The
// result is the correctly rounded field of the fixed-point couplings: exact for integer J, and
// for float J more accurate (2^-31 of max|J| per coupling) than any fp32 summation order.
// (bf16 planes with fp32 accumulation would not do here: the tensor core adds into the fp32
// accumulator with truncation, a bias of ~0.35 ulp per MMA that 768 accumulation steps turn into
// 3-6e-5 -- measured on the sweep kernel's drift, profiles/r1_notes.md.)
//
// Shapes: CTA = 128 field columns (one M tile) x 128 replicas (N) x 4 digit accumulators = the
// whole 512-column TMEM; K loop over blocks of 32 spins (one MMA K step).  Operands are pre-tiled
// in HBM in the UMMA canonical no-swizzle layouts (digits: static per model; spins: per call) so
// that every K step is two TMA bulk copies (16 KB of digits + 4 KB of spins) into an 8-stage ring.
#include <cstdint>

#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_tc.cuh"

namespace sg {

namespace {

constexpr int kFM = 128;                 // field columns per CTA (MMA M)
constexpr int kFN = 128;                 // replicas per CTA (MMA N)
constexpr int kFK = 32;                  // spins per K step (int8 MMA K)
constexpr int kDigits = 4;
constexpr int kATile = kFM * kFK;        // 4096 B per digit
constexpr int kBTile = kFN * kFK;        // 4096 B
constexpr int kFStage = kDigits * kATile + kBTile;  // 20 KB
constexpr int kFStages = 8;
constexpr uint32_t kFA_LBO = 1024, kFA_SBO = 128;   // A (MN-major, 8-bit): k-group, m-group stride
constexpr uint32_t kFB_LBO = 128, kFB_SBO = 256;    // B (K-major, 8-bit): k-group, n-group stride
// kind::i8: D = s32, A = B = signed 8 bit, A MN-major, B K-major
constexpr uint32_t kFIdesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) |
                             ((uint32_t)(kFN >> 3) << 17) | ((uint32_t)(kFM >> 4) << 24);

__device__ __forceinline__ void mma_i8_ss(uint32_t taddr_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(taddr_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---------------------------------------------------------------- model: fixed-point digits
// info[0] = max |J| (float bits), info[1] = 1 if some coupling is not an integer
__global__ void absmax_kernel(const float* __restrict__ Jt, int n, int n_pad, unsigned int* info) {
    unsigned int m = 0, frac = 0;
    const size_t total = (size_t)n * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / n), j = (int)(idx - (size_t)i * n);
        const float x = Jt[(size_t)i * n_pad + j];
        m = max(m, __float_as_uint(fabsf(x)));
        frac |= (x != rintf(x)) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        frac |= __shfl_xor_sync(0xFFFFFFFFu, frac, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&info[0], m);
        if (frac) atomicOr(&info[1], 1u);
    }
}

// scale[0] = 2^s, scale[1] = 2^-s
__global__ void scale_kernel(const unsigned int* info, double* scale) {
    const float jmax = __uint_as_float(info[0]);
    int s = 0;
    if (jmax > 0.0f && (info[1] != 0u || jmax >= 1073741824.0f)) {
        int e;
        frexpf(jmax, &e);  // jmax = f * 2^e, f in [0.5, 1)  =>  jmax * 2^(30 - e) < 2^30
        s = 30 - e;
    }
    scale[0] = ldexp(1.0, s);
    scale[1] = ldexp(1.0, -s);
}

// digits[kb][jt][d][tile]: byte(m, k) = (m % 16) + (k % 8) * 16 + (m / 16) * 128 + (k / 8) * 1024
// one thread = one (i, 16 consecutive j): 64 B read, 4 x 16 B written
__global__ void digits_kernel(const float* __restrict__ Jt, int n, int n_pad, int n_tc,
                              const double* __restrict__ scale, unsigned char* __restrict__ dig) {
    const int njt = n_tc / kFM;
    const int nkb = (n + kFK - 1) / kFK;
    const size_t total = (size_t)nkb * kFK * (n_tc / 16);
    const double sc = scale[0];
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int jg = (int)(idx % (n_tc / 16));   // group of 16 columns
        const int i = (int)(idx / (n_tc / 16));
        const int j0 = jg * 16;
        int a[4][16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int j = j0 + e;
            const float x = (i < n && j < n) ? Jt[(size_t)i * n_pad + j] : 0.0f;
            long long q = llrint((double)x * sc);
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int ad = (int)(((q + 128) & 255) - 128);
                a[d][e] = ad;
                q = (q - ad) >> 8;
            }
        }
        const int kb = i / kFK, k = i % kFK;
        const int jt = j0 / kFM, m0 = j0 % kFM;
        unsigned char* tile = dig + (((size_t)kb * njt + jt) * kDigits) * kATile;
        const int off = (k & 7) * 16 + (m0 >> 4) * 128 + (k >> 3) * 1024;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            uint32_t w[4];
#pragma unroll
            for (int v = 0; v < 4; ++v)
                w[v] = (uint32_t)(a[d][4 * v] & 255) | ((uint32_t)(a[d][4 * v + 1] & 255) << 8) |
                       ((uint32_t)(a[d][4 * v + 2] & 255) << 16) |
                       ((uint32_t)(a[d][4 * v + 3] & 255) << 24);
            *reinterpret_cast<uint4*>(tile + (size_t)d * kATile + off) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// spins [R][ld] int8 -> tiles[rt][kb]: byte(r, k) = (r % 8) * 16 + (k % 16) + (r / 8) * 256 + (k / 16) * 128
__global__ void spin_tiles_kernel(const int8_t* __restrict__ S, int64_t ld, int n, int R,
                                  unsigned char* __restrict__ tiles) {
    const int nkb = (n + kFK - 1) / kFK;
    const int nrt = (R + kFN - 1) / kFN;
    const size_t total = (size_t)nrt * kFN * nkb * 2;  // 16-byte units
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int kh = (int)(idx % (nkb * 2));       // half K block: 16 spins
        const int r = (int)(idx / (nkb * 2));
        const int i0 = kh * 16;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < R) {
            if (i0 + 16 <= n) {
                v = *reinterpret_cast<const uint4*>(S + (size_t)r * ld + i0);
            } else {
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                for (int e = 0; e < 16; ++e)
                    if (i0 + e < n) w[e >> 2] |= (uint32_t)(uint8_t)S[(size_t)r * ld + i0 + e] << (8 * (e & 3));
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        const int rt = r / kFN, rr = r % kFN, kb = kh >> 1;
        unsigned char* tile = tiles + ((size_t)rt * nkb + kb) * kBTile;
        *reinterpret_cast<uint4*>(tile + (rr & 7) * 16 + (rr >> 3) * 256 + (kh & 1) * 128) = v;
    }
}

// ---------------------------------------------------------------- the GEMM
__global__ void __launch_bounds__(192, 1)
fields_tc_kernel(const unsigned char* __restrict__ dig, const unsigned char* __restrict__ stiles,
                 const float* __restrict__ h, const double* __restrict__ scale, int n, int n_tc,
                 int R, float* __restrict__ F, int64_t ldF) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kFStages * kFStage);
    uint64_t* empty = full + kFStages;
    uint64_t* done = empty + kFStages;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int jt = blockIdx.x, rt = blockIdx.y;
    const int njt = n_tc / kFM;
    const int nkb = (n + kFK - 1) / kFK;

    if (tid == 0) {
        for (int d = 0; d < kFStages; ++d) {
            mbar_init(&full[d], 1);
            mbar_init(&empty[d], 1);
        }
        mbar_init(done, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == 0) {
        tc::tmem_alloc(tptr, 512);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *tptr;

    if (warp == 4) {
        // ---- producer: digits (16 KB) + spins (4 KB) of K block kb
        if (lane == 0) {
            int stage = 0;
            uint32_t epar = 1;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&empty[stage], epar);
                mbar_arrive_expect_tx(&full[stage], (uint32_t)kFStage);
                unsigned char* dst = ring + (size_t)stage * kFStage;
                bulk_g2s(dst, dig + (((size_t)kb * njt + jt) * kDigits) * kATile,
                         (uint32_t)(kDigits * kATile), &full[stage]);
                bulk_g2s(dst + kDigits * kATile, stiles + ((size_t)rt * nkb + kb) * kBTile,
                         (uint32_t)kBTile, &full[stage]);
                if (++stage == kFStages) { stage = 0; epar ^= 1u; }
            }
        }
    } else if (warp == 5) {
        // ---- MMA issuer: D_d += A_d * B for the four digits
        int stage = 0;
        uint32_t fpar = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full[stage], fpar);
            tc::fence_after_sync();
            const uint32_t sbase = smem_u32(ring + (size_t)stage * kFStage);
            const uint64_t adesc0 = tc::make_smem_desc(sbase, kFA_LBO, kFA_SBO);
            const uint64_t bdesc = tc::make_smem_desc(sbase + kDigits * kATile, kFB_LBO, kFB_SBO);
            if (tc::elect_one()) {
#pragma unroll
                for (int d = 0; d < kDigits; ++d)
                    mma_i8_ss(tbase + d * kFN, adesc0 + (uint64_t)((d * kATile) >> 4), bdesc, kFIdesc,
                              kb > 0 ? 1u : 0u);
                tc::mma_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == kFStages) { stage = 0; fpar ^= 1u; }
        }
        if (tc::elect_one()) tc::mma_commit(done);
        __syncwarp();
    } else {
        // ---- epilogue (warps 0-3 = TMEM lane quarters): recombine digits, scale, add h, store
        mbar_wait(done, 0);
        tc::fence_after_sync();
        const uint32_t tq = tbase + ((uint32_t)(warp * 32) << 16);
        const int j = jt * kFM + warp * 32 + lane;
        const bool j_ok = j < ldF;          // the row length n_tc may exceed the padded row of F / h
        const float hv = j_ok ? h[j] : 0.0f;
        const double inv = scale[1];
#pragma unroll 1
        for (int c = 0; c < kFN / 16; ++c) {
            uint32_t d0[16], d1[16], d2[16], d3[16];
            tc::tmem_ld16_u32(tq + 0 * kFN + c * 16, d0);
            tc::tmem_ld16_u32(tq + 1 * kFN + c * 16, d1);
            tc::tmem_ld16_u32(tq + 2 * kFN + c * 16, d2);
            tc::tmem_ld16_u32(tq + 3 * kFN + c * 16, d3);
            tc::wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int r = rt * kFN + c * 16 + e;
                const long long q = (long long)(int)d0[e] + ((long long)(int)d1[e] << 8) +
                                    ((long long)(int)d2[e] << 16) + ((long long)(int)d3[e] << 24);
                if (r < R && j_ok) F[(size_t)r * ldF + j] = (float)((double)q * inv + (double)hv);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

}  // namespace

size_t fields_tc_digits_bytes(int n, int n_tc) {
    const int nkb = (n + kFK - 1) / kFK;
    return (size_t)nkb * (n_tc / kFM) * kDigits * kATile;
}

size_t fields_tc_spin_tiles_bytes(int n, int R) {
    const int nkb = (n + kFK - 1) / kFK;
    return (size_t)((R + kFN - 1) / kFN) * nkb * kBTile;
}

// once per model: info = 2 x u32 scratch, scale = 2 doubles, dig = fields_tc_digits_bytes()
cudaError_t launch_fields_tc_prepare(const float* Jt, int n, int n_pad, int n_tc, unsigned int* info,
                                     double* scale, void* dig, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(info, 0, 2 * sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    absmax_kernel<<<592, 256, 0, st>>>(Jt, n, n_pad, info);
    scale_kernel<<<1, 1, 0, st>>>(info, scale);
    digits_kernel<<<1184, 256, 0, st>>>(Jt, n, n_pad, n_tc, scale, static_cast<unsigned char*>(dig));
    return cudaGetLastError();
}

// F[r][0..n_tc) = S J^T + h for R configurations (columns n..n_tc come out as h = 0)
cudaError_t launch_fields_tc(const int8_t* spins, int64_t ld_spins, const void* dig,
                             const double* scale, const float* h, int n, int n_tc, int R,
                             void* spin_tiles, float* fields, int64_t ld_fields, cudaStream_t st) {
    {
        const size_t units = fields_tc_spin_tiles_bytes(n, R) / 16;
        int grid = (int)((units + 255) / 256 < (size_t)2368 ? (units + 255) / 256 : (size_t)2368);
        spin_tiles_kernel<<<grid, 256, 0, st>>>(spins, ld_spins, n, R,
                                                static_cast<unsigned char*>(spin_tiles));
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const int smem = kFStages * kFStage + (2 * kFStages + 1) * 8 + 16;
    cudaError_t e = cudaFuncSetAttribute(fields_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    dim3 grid(n_tc / kFM, (R + kFN - 1) / kFN);
    fields_tc_kernel<<<grid, 192, smem, st>>>(static_cast<const unsigned char*>(dig),
                                              static_cast<const unsigned char*>(spin_tiles), h, scale,
                                              n, n_tc, R, fields, ld_fields);
    return cudaGetLastError();
}

}  // namespace sg
