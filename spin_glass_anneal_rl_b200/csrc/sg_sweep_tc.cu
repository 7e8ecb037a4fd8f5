// sg_sweep_tc.cu -- K1-TC: the replica-batched Monte Carlo sweep on the 5th-gen tensor cores.
//
// Same contract as sg_sweep.cu (SpinDynamics.sweep() for R replicas, reference
// core/spin_dynamics.py:61-94,131-191), different engine.  The local fields
// f[r][j] = h_j + sum_i J_ji s_ri of the block's 16 replicas are fp32 ACCUMULATORS RESIDENT IN
// TENSOR MEMORY (16 replicas x 4096 columns x 4 B = the whole 256 KB TMEM of an SM: field
// column j of replica r is TMEM lane j % 128, column (j / 128) * 16 + r).  The sequential
// algorithm is processed in blocks of 16 consecutive attempts.  For a block the accepted flips
// are a 16 x 16 matrix D (attempt x replica, entries -2 s or 0) and the field update of the
// whole block is the rank-16 product
//
//        F[:, r] += sum_k Jt[site_k][:] * D[k][r]          (F: 4096 x 16, Jt rows: 4096 x 16)
//
// i.e. 32 tcgen05.mma (M = 128 field columns, N = 16 replicas, K = 16 attempts) per bf16 plane
// of J.  J is held as P bf16 planes whose sum is the coupling (P = 3 reproduces every fp32
// coupling exactly: 3 x 8 mantissa bits; P = 2 keeps 16 bits, P = 1 is plain bf16); the
// products (+-2 x bf16) are exact and the accumulation is fp32, so for integer couplings the
// fields -- and with them every accept decision -- are exact.
//
// Because the update of a block is applied only after the block has been decided, the decision
// warp corrects the 16 field values it needs itself, exactly as in sg_sweep.cu: raw values as
// of two blocks ago + a 16x16 table of couplings from the previous block's sites + the in-block
// table, with the same FMAs in the same order as the sequential algorithm.
//
// Warp roles (10 warps, 1 block per SM):
//   warps 0-3  "quarter" warps (TMEM lane quarter = warp id): read the raw field values of the
//              16 sites of block k+2 from TMEM (tcgen05.ld), gather the coupling tables, draw the
//              Philox thresholds; at sweep end reduce the energies from TMEM;
//   warps 4-7  producers: gather the 16 J rows of a block from L2 into the UMMA canonical
//              operand layout with 16-byte cp.async (LDGSTS), chunk by chunk (4 tiles x P planes)
//              through a shared-memory ring; completion is signalled on mbarriers;
//   warp 8     decision warp (lane = replica): 16 attempts per block in registers, writes the
//              B operand (bf16 deltas), flips the spin bit planes;
//   warp 9     MMA issuer (one lane): tcgen05.mma + tcgen05.commit.
#include <cuda_bf16.h>

#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_tc.cuh"

namespace sg {

namespace {

constexpr int kG = 16;             // replicas per block = MMA N
constexpr int kTileM = 128;        // field columns per MMA
constexpr int kBlk = 16;           // attempts per block = MMA K
constexpr int kTileBytes = kTileM * kBlk * 2;   // one (tile, plane) A operand: 4096 B
constexpr int kChunkTiles = 4;     // tiles per ring stage
constexpr uint32_t kIdesc = tc::make_idesc_bf16(kTileM, kG);
constexpr uint32_t kALbo = 2048, kASbo = 128;   // A: k-group stride, m-group stride
constexpr uint32_t kBLbo = 256, kBSbo = 128;    // B: k-group stride, n-group stride
constexpr int kBopBytes = 512;

// ---------------------------------------------------------------- model planes
// Jp[p][i][j] (bf16): Jt[i][j] = sum_p Jp[p][i][j] (+ residual below 2^-24 relative for p = 3)
__global__ void split_planes_kernel(const float* __restrict__ Jt, int n, int n_pad,
                                    __nv_bfloat16* __restrict__ Jp, int n_tc) {
    const size_t total = (size_t)n * n_tc;
    const size_t plane = total;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / n_tc), j = (int)(idx - (size_t)i * n_tc);
        const float x = (j < n) ? Jt[(size_t)i * n_pad + j] : 0.0f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(x);
        const float r1 = x - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
        const float r2 = r1 - __bfloat162float(mid);
        const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
        Jp[idx] = hi;
        Jp[plane + idx] = mid;
        Jp[2 * plane + idx] = lo;
    }
}

// ---------------------------------------------------------------- operand gather
// One producer thread's share of a chunk.  The 128 producer threads are 4 warps `pw`; thread
// (pw, lane) owns k-row kk = lane % 8 of both k-groups (sites site[0], site[1]) and m-group
// mg = 4 * pw + lane / 8 of every tile:
//   dst(tile tt, plane p)  = chunk + (tt * P + p) * 4096
//   within a tile: byte(m, k) = (m/8)*128 + (k/8)*2048 + (k%8)*16 + (m%8)*2
// so one warp instruction moves 8 k-rows x 4 m-groups = 512 contiguous bytes of shared memory
// (conflict-free) from 8 row segments of 64 B.
template <int P>
__device__ __forceinline__ void gather_chunk(unsigned char* chunk, const __nv_bfloat16* Jp,
                                             size_t plane_stride, int n_tc, int tile0, int ntiles,
                                             const int (&site)[2], int pw, int lane) {
    const int kk = lane & 7, mg = pw * 4 + (lane >> 3);
#pragma unroll
    for (int p = 0; p < P; ++p) {
#pragma unroll
        for (int kg = 0; kg < 2; ++kg) {
            const __nv_bfloat16* row =
                Jp + (size_t)p * plane_stride + (size_t)site[kg] * n_tc + mg * 8;
            unsigned char* dst = chunk + p * kTileBytes + mg * 128 + kg * 2048 + kk * 16;
#pragma unroll
            for (int tt = 0; tt < kChunkTiles; ++tt)
                if (tt < ntiles)
                    tc::cp_async16(dst + tt * P * kTileBytes, row + (size_t)(tile0 + tt) * kTileM);
        }
    }
}

// ---------------------------------------------------------------- self test
// One block: fields (16 x n_tc) -> TMEM, one rank-16 update with the given sites / deltas,
// TMEM -> out.  Validates the operand layouts, descriptors and the TMEM addressing.
template <int P>
__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const __nv_bfloat16* __restrict__ Jp, int n, int n_tc,
                   const int* __restrict__ sites, const float* __restrict__ deltas,
                   const float* __restrict__ fields_in, float* __restrict__ fields_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* chunk = smem;                               // 4 * P * 4096
    unsigned char* bop = smem + kChunkTiles * P * kTileBytes;  // 512
    uint64_t* bar = reinterpret_cast<uint64_t*>(bop + kBopBytes);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntile = n_tc / kTileM;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tc::tmem_alloc(tptr, 512);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = *tptr;
    const uint32_t tq = tbase + ((uint32_t)(warp * 32) << 16);

    // fields -> TMEM
    for (int t = 0; t < ntile; ++t) {
        float v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = fields_in[(size_t)r * n_tc + t * kTileM + warp * 32 + lane];
        tc::tmem_st16(tq + t * kG, v);
    }
    tc::wait_st();
    // B operand
    if (tid < 16) {
        const int r = tid;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const __nv_bfloat16 d = __float2bfloat16_rn(deltas[k * 16 + r]);
            *reinterpret_cast<__nv_bfloat16*>(bop + (r & 7) * 16 + (r >> 3) * kBSbo + (k & 7) * 2 +
                                              (k >> 3) * kBLbo) = d;
        }
    }
    int site[2] = {sites[(lane & 7)], sites[8 + (lane & 7)]};
    tc::fence_before_sync();
    __syncthreads();

    uint32_t parity = 0;
    for (int tile0 = 0; tile0 < ntile; tile0 += kChunkTiles) {
        const int nt = min(kChunkTiles, ntile - tile0);
        gather_chunk<P>(chunk, Jp, (size_t)n * n_tc, n_tc, tile0, nt, site, warp, lane);
        tc::cp_async_wait_all();
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            const uint64_t bdesc = tc::make_smem_desc(smem_u32(bop), kBLbo, kBSbo);
            for (int tt = 0; tt < nt; ++tt)
                for (int p = 0; p < P; ++p) {
                    const uint64_t adesc = tc::make_smem_desc(
                        smem_u32(chunk + (tt * P + p) * kTileBytes), kALbo, kASbo);
                    tc::mma_bf16_ss(tbase + (tile0 + tt) * kG, adesc, bdesc, kIdesc, 1u);
                }
            tc::mma_commit(&bar[0]);
        }
        mbar_wait(&bar[0], parity);
        parity ^= 1u;
        tc::fence_after_sync();
        __syncthreads();
    }

    // TMEM -> out
    for (int t = 0; t < ntile; ++t) {
        float v[16];
        tc::tmem_ld16(tq + t * kG, v);
        tc::wait_ld();
#pragma unroll
        for (int r = 0; r < 16; ++r)
            fields_out[(size_t)r * n_tc + t * kTileM + warp * 32 + lane] = v[r];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

}  // namespace

cudaError_t launch_split_planes(const float* Jt, int n, int n_pad, void* Jp, int n_tc,
                                cudaStream_t st) {
    split_planes_kernel<<<1184, 256, 0, st>>>(Jt, n, n_pad, static_cast<__nv_bfloat16*>(Jp), n_tc);
    return cudaGetLastError();
}

cudaError_t launch_tc_selftest(const void* Jp, int n, int n_tc, int planes, const int* sites,
                               const float* deltas, const float* fields_in, float* fields_out,
                               cudaStream_t st) {
    const __nv_bfloat16* J = static_cast<const __nv_bfloat16*>(Jp);
    cudaError_t err;
#define SG_ST(P)                                                                                 \
    {                                                                                            \
        const int smem = kChunkTiles * P * kTileBytes + kBopBytes + 64;                          \
        err = cudaFuncSetAttribute(tc_selftest_kernel<P>,                                        \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, smem);           \
        if (err != cudaSuccess) return err;                                                      \
        tc_selftest_kernel<P><<<1, 128, smem, st>>>(J, n, n_tc, sites, deltas, fields_in,        \
                                                    fields_out);                                 \
    }
    if (planes == 1) SG_ST(1) else if (planes == 2) SG_ST(2) else if (planes == 3) SG_ST(3)
    else return cudaErrorInvalidValue;
#undef SG_ST
    return cudaGetLastError();
}

}  // namespace sg
