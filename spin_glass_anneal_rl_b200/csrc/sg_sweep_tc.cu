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
// The J rows of a block must reach shared memory in the UMMA canonical operand layout
// (A is MN-major: 8 attempts x 8 columns core matrices), which is a 16-byte-granular transpose
// of 16 randomly chosen rows.  Doing that gather inside the sweep kernel with cp.async (LDGSTS)
// tops out near 13 B/clk/SM and floods the LSU queue the decision warp depends on (measured),
// so the gather is done ONCE PER SWEEP FOR THE WHOLE GRID by a separate streaming kernel
// (tc_gather_kernel: every block visits the sites in the same order, so all 148 SMs consume the
// same operand stream) and the sweep kernel pulls ready-made 4-tile chunks through a
// shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier), L2-resident for all
// blocks but the first that touches a chunk.
//
// Warp roles (7 warps, 1 block per SM):
//   warps 0-3  "quarter" warps (TMEM lane quarter = warp id): read the raw field values of the
//              16 sites of block k+2 from TMEM (tcgen05.ld), gather the coupling tables, draw the
//              Philox thresholds; at sweep end reduce the energies from TMEM;
//   warp 4     producer (one lane): TMA bulk copies of operand chunks into the ring;
//   warp 5     decision warp (lane = replica): 16 attempts per block in registers, writes the
//              B operand (bf16 deltas), flips the spin bit planes;
//   warp 6     MMA issuer (one lane): tcgen05.mma + tcgen05.commit.
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_tc.cuh"

namespace sg {

namespace {

constexpr int kG = 16;             // replicas per block = MMA N
constexpr int kTileM = 128;        // field columns per MMA
constexpr int kBlk = 16;           // attempts per block = MMA K
// A operand: m-group (core matrix) stride.  Core matrices need not be 128-byte aligned (144 was
// tried to spread a k-slice over all banks: bit-exact results, not faster), so keep them dense.
constexpr uint32_t kASbo = 128;
constexpr uint32_t kALbo = 16 * kASbo;          // A: k-group stride
constexpr int kTileBytes = 2 * kALbo;           // one (tile, plane) A operand
constexpr int kChunkTiles = 4;     // tiles per ring stage
constexpr uint32_t kIdesc = tc::make_idesc_bf16(kTileM, kG);
constexpr uint32_t kBLbo = 256, kBSbo = 128;    // B: k-group stride, n-group stride
constexpr int kBopBytes = 512;

// ---------------------------------------------------------------- model planes
// Jp[p][i][j] (bf16): Jt[i][j] = sum_p Jp[p][i][j] (+ residual below 2^-24 relative for p = 3)
// used[0] / used[1] are set when the second / third plane holds a non-zero entry: models whose
// couplings fit fewer planes exactly (integer couplings: one) need fewer MMAs per block
__global__ void split_planes_kernel(const float* __restrict__ Jt, int n, int n_pad,
                                    __nv_bfloat16* __restrict__ Jp, int n_tc, int* __restrict__ used) {
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
        if (used) {
            if (r1 != 0.0f) used[0] = 1;
            if (r2 != 0.0f) used[1] = 1;
        }
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
            unsigned char* dst = chunk + p * kTileBytes + mg * kASbo + kg * kALbo + kk * 16;
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


// ---------------------------------------------------------------- site tables
// sites[s][i] (uint16, row length n_s = n rounded up to 16, padding = site 0) for every sweep of
// a launch; one order for the whole grid (the J rows a block gathers are then L2 hits for all
// blocks but the first).  Same Philox stream as gen_sites() of sg_sweep.cu.
__global__ void tc_sites_kernel(int mode, unsigned long long seed, unsigned long long sweep_base,
                                int n, int n_s, int n_sweeps, const int* __restrict__ explicit_sites,
                                long long s_ss, uint16_t* __restrict__ out) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const int quads = n_s / 4;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_sweeps * quads;
         idx += gridDim.x * blockDim.x) {
        const int s = idx / quads, q = idx - s * quads;
        uint32_t v[4] = {0u, 0u, 0u, 0u};
        if (mode == 1) {
            const unsigned long long sa = sweep_base + (unsigned long long)s;
            const uint4 x = philox4x32_10(
                make_uint4(kSiteStreamTag, (uint32_t)sa, (uint32_t)(sa >> 32), (uint32_t)q), key);
            v[0] = x.x % (uint32_t)n; v[1] = x.y % (uint32_t)n;
            v[2] = x.z % (uint32_t)n; v[3] = x.w % (uint32_t)n;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = q * 4 + e;
            uint32_t site = 0;
            if (i < n) {
                if (mode == 0) site = (uint32_t)i;
                else if (mode == 1) site = v[e];
                else site = (uint32_t)explicit_sites[(long long)s * s_ss + i];
            }
            out[(size_t)s * n_s + i] = (uint16_t)site;
        }
    }
}


// ---------------------------------------------------------------- operand stream
// Q[(s, block k, chunk c)] = kStage bytes = 4 tiles x P planes, each tile the 128 x 16 A operand
// of one tcgen05.mma in canonical layout: byte(m, k) = (m/8)*kASbo + (k/8)*kALbo + (k%8)*16 +
// (m%8)*2, value Jp[p][site_k][tile*128 + m].  One thread moves one 16-byte unit; consecutive
// threads write consecutive units (coalesced) and read 8 rows x 64 B.
template <int P>
__global__ void __launch_bounds__(256)
tc_gather_kernel(const __nv_bfloat16* __restrict__ Jp, int n, int n_tc,
                 const uint16_t* __restrict__ sites, int n_s, int s_begin, int n_sub, int nblk,
                 int nchunk, uint4* __restrict__ Q) {
    constexpr int kUnitsPerTile = kTileBytes / 16;
    constexpr int kUnitsPerStage = kChunkTiles * P * kUnitsPerTile;
    const size_t plane_stride = (size_t)n * n_tc;
    const int T = n_tc / kTileM;
    const size_t total = (size_t)n_sub * nblk * nchunk * kUnitsPerStage;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (size_t)gridDim.x * blockDim.x) {
        const int u = (int)(g % kUnitsPerStage);
        const size_t chunk_id = g / kUnitsPerStage;
        const int c = (int)(chunk_id % nchunk);
        const size_t blk_id = chunk_id / nchunk;
        const int k = (int)(blk_id % nblk);
        const int s_rel = (int)(blk_id / nblk);
        const int tp = u / kUnitsPerTile, d = u - tp * kUnitsPerTile;
        const int tt = tp / P, p = tp - tt * P;
        // d*16 = mg*kASbo + kg*kALbo + kk*16  (kASbo = 128 or 144, kALbo = 16*kASbo)
        const int off = d * 16;
        const int kg = off / (int)kALbo, rem = off - kg * (int)kALbo;
        const int mg = rem / (int)kASbo, rem2 = rem - mg * (int)kASbo;
        const int kk = rem2 >> 4;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        const int tile = c * kChunkTiles + tt;
        if (tile < T && kk < 8) {
            const int site = sites[(size_t)(s_begin + s_rel) * n_s + k * kBlk + kg * 8 + kk];
            v = *reinterpret_cast<const uint4*>(Jp + (size_t)p * plane_stride + (size_t)site * n_tc +
                                                tile * kTileM + mg * 8);
        }
        Q[g] = v;
    }
}

// Decision tables of every attempt block, computed once per sweep for the whole grid:
//   cin[a][b]  = J'[site_a][site_b]                (couplings among the block's own sites)
//   ccr1[a][b] = J'[site_a of block k-1][site_b]   (from the previous block's sites; 0 for k = 0)
//   ccr2[a][b] = J'[site_a of block k-2][site_b]   (two blocks back; 0 for k < 2; cluster variant)
//   dup[b]     = bit mask of the earlier attempts of the block that visit the same site
// J' = sum of the P planes.  kTabBytes per block, fetched by the sweep kernel with one TMA copy.
constexpr int kTabBytes = 3 * kBlk * kBlk * 4 + kBlk * 4;  // 3136

template <int P>
__global__ void __launch_bounds__(256)
tc_tables_kernel(const __nv_bfloat16* __restrict__ Jp, int n, int n_tc,
                 const uint16_t* __restrict__ sites, int n_s, int s_begin, int nblk,
                 unsigned char* __restrict__ tabs) {
    const int k = blockIdx.x % nblk, s_rel = blockIdx.x / nblk;
    const uint16_t* stab = sites + (size_t)(s_begin + s_rel) * n_s;
    const int i0 = k * kBlk;
    const int nbk = min(kBlk, n - i0);
    const size_t plane_stride = (size_t)n * n_tc;
    unsigned char* out = tabs + (size_t)blockIdx.x * kTabBytes;
    float* cin = reinterpret_cast<float*>(out);
    uint32_t* dup = reinterpret_cast<uint32_t*>(cin + 3 * kBlk * kBlk);
    for (int idx = threadIdx.x; idx < 3 * kBlk * kBlk; idx += blockDim.x) {
        const int which = idx >> 8, aa = (idx >> 4) & 15, b = idx & 15;
        float val = 0.0f;
        if (b < nbk && (which == 0 ? (aa < nbk) : (k >= which))) {
            const int sb = stab[i0 + b];
            const int sr = stab[i0 - which * kBlk + aa];
            const __nv_bfloat16* pj = Jp + (size_t)sr * n_tc + sb;
            val = __bfloat162float(pj[0]);
            if (P > 1) val += __bfloat162float(pj[plane_stride]);
            if (P > 2) val += __bfloat162float(pj[2 * plane_stride]);
        }
        cin[which * kBlk * kBlk + aa * kBlk + b] = val;
    }
    if (threadIdx.x < kBlk) {
        uint32_t m = 0;
        const int me = stab[i0 + threadIdx.x];
        for (int a2 = 0; a2 < (int)threadIdx.x; ++a2) m |= (stab[i0 + a2] == me) ? (1u << a2) : 0u;
        dup[threadIdx.x] = m;
    }
}

// ---------------------------------------------------------------- the sweep
constexpr int kSlots = 4;
// Column parts of a CTA's tiles: the raw reads of block k+1 in a part chase the MMA wavefront of
// block k-1 and must be done before block k's MMAs of that part are issued.  Two parts; four were
// tried (the read of one part would overlap the execution of three others) and were slower on
// every cluster size (C = 4: 16.6 instead of 19.2 G attempts/s): every part costs the MMA warp a
// barrier round trip through the quarter warps.  Also tried for C = 4 (profiles/r2_notes.md):
// issuing the tiles that hold no site of the next block first (they need no read) and the others
// after the next block's reads -- correct, but a ring stage is then released only with its last
// tile, the five-stage ring (1.25 blocks) cannot prefetch the next block and the TMA latency lands
// on the critical path (14.8 G); releasing the ring stages through the part commits instead of a
// commit per stage (no gain, and the producer can fall two barrier phases behind).  Also tried for C = 4: issuing the tiles that hold
// no site of the next block first (they need no read) and the others after the next block's reads
// -- correct, but a stage of the ring is then released only with its last tile, the five-stage
// ring (1.25 blocks) cannot prefetch the next block, and the TMA latency lands on the critical
// path (14.8 G attempts/s).  It would need a ring of two blocks, i.e. 70 KB more shared memory.
constexpr int kParts = 2;
constexpr int kMaxStagesTc = 8;
// C = CTAs per replica group: NG = 16 C replicas, decision warps = one per 32 replicas
__host__ __device__ constexpr int tc_ndw(int C) { return (kG * C + 31) / 32; }
// threshold warps: clusters of 4 and more draw the Philox thresholds on two dedicated warps, so that
// the quarter warps sit on the half-done barriers and answer a completed half at once (the round
// trip MMA complete -> raw read -> next MMA issue is what bounds the block period there)
__host__ __device__ constexpr int tc_ntw(int C) { return C >= 4 ? 2 : 0; }
// threads: 4 quarter warps, producer, decision warp 0, MMA issuer (, decision warps 1..3)
// (, threshold warps)
// Warp slots (warps are dealt to the four SM sub-partitions round robin, slot % 4): the two
// Philox-heavy threshold warps take slots 4 and 6 + ndw (8 for clusters of 4), both on sub-partition 0,
// and the TMA producer -- one lane, next to no instructions -- moves to the slot after them (9,
// sub-partition 1, next to decision warp 0).  Measured alternatives: threshold warps on slots 8 and 9
// slow decision warp 0, whose instruction stream is the critical path (16 attempts in 1.44 k clocks
// against 1.15 k undisturbed); on slots 8 and 10 the second one slows the MMA issuer; a single
// threshold warp cannot keep up (14.8 G attempts/s).
__host__ __device__ constexpr int tc_threads(int C) {
    return tc_ntw(C) ? 32 * (6 + tc_ndw(C) + 2) : 224 + 32 * (tc_ndw(C) - 1);
}
// quarter warps + decision warps meet at the named barriers 1..3
__host__ __device__ constexpr int tc_sync_threads(int C) { return 128 + 32 * tc_ndw(C); }
// tiles per ring stage: clusters of 4 hold 64 replicas' state in shared memory and have room for
// 120 KB of ring only -- five 2-tile stages keep more copies in flight than two 4-tile ones
__host__ __device__ constexpr int tc_chunk_tiles(int C) { return C == 8 ? 1 : C == 4 ? 2 : kChunkTiles; }

struct TcSmem {
    size_t ring, bop, sbits, theta, raw, tab, ztab, red, flags, bars, tptr, total;
};

// C = CTAs per replica group (1, or 2 = a cluster pair that splits the field columns)
__host__ __device__ inline TcSmem tc_layout(int n_tc, int P, int NS, int C) {
    const int NG = kG * C;
    TcSmem L;
    size_t off = 0;
    L.ring = off;  off += (size_t)NS * tc_chunk_tiles(C) * P * kTileBytes;
    L.bop = off;   off += (size_t)kSlots * 32 * NG;
    L.sbits = off; off += (size_t)NG * (n_tc / 32 + 1) * sizeof(uint32_t);   // stride W + 1: no bank conflicts
    L.theta = off; off += (size_t)kSlots * kBlk * NG * sizeof(float);
    L.raw = off;   off += (size_t)kSlots * kBlk * NG * sizeof(float);
    L.tab = off;   off += (size_t)kSlots * kTabBytes;
    L.ztab = off;  off += (size_t)kBlk * kBlk * sizeof(float);   // a cross table of zeros
    L.red = off;   off += (size_t)C * 4 * NG * sizeof(float);
    L.flags = off; off += 4 * sizeof(uint32_t);
    off = (off + 15) & ~(size_t)15;
    L.bars = off;  off += (size_t)(2 * kMaxStagesTc + (5 + 2 * kParts) * kSlots + 2) * sizeof(uint64_t);
    L.tptr = off;  off += 16;
    L.total = off;
    return L;
}

template <int C>
__device__ __forceinline__ void named_sync_c(int id) { named_bar_sync(id, tc_sync_threads(C)); }

// ---- thread-block-cluster helpers (C = 2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in the CTA of rank `rank`
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// asynchronous remote store that counts its bytes on an mbarrier of the target CTA (both
// addresses are shared::cluster addresses of the same CTA): no fence, no remote arrive
__device__ __forceinline__ void st_async_f4(uint32_t addr, float4 v, uint32_t bar) {
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
        "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void st_async_f2(uint32_t addr, float2 v, uint32_t bar) {
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr),
        "f"(v.x), "f"(v.y), "r"(bar)
        : "memory");
}

// development aid: clock stamps of block 0 (tools/tc_timeline.py); 16 slots per attempt block
#define SG_ISTAMP(slot_)                                                             \
    do {                                                                             \
        if (a.dbg && blockIdx.x == 0 && tid == 0 && it == cid) a.dbg[10240 + (slot_)] = clock64(); \
    } while (0)
#define SG_STAMP(slot_)                                                              \
    do {                                                                             \
        if (a.dbg && blockIdx.x == 0 && lane == 0 && kg < 512) a.dbg[kg * 16 + (slot_)] = clock64(); \
    } while (0)

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// 32 spin bits -> 32 int8 (+1 / -1), written as two 16-byte stores
__device__ __forceinline__ void store_spin_word(int8_t* dst, uint32_t w) {
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t nib = (w >> (4 * j)) & 0xFu;
        uint32_t bytes = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) bytes |= (((nib >> e) & 1u) ? 0x01u : 0xFFu) << (8 * e);
        o[j] = bytes;
    }
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
    d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// Persistent-CTA work list of the sweep kernel.
struct TcItems {
    int n_groups, s_begin, s_end, spi, n_cta;
    __host__ __device__ int n_chunks() const { return (s_end - s_begin + spi - 1) / spi; }
    __host__ __device__ int count() const { return n_groups * n_chunks(); }
};
struct TcItem {
    int ch, g, s_lo, s_hi;
};
__host__ __device__ inline TcItem tc_item(const TcItems& w, int it) {
    TcItem x;
    x.ch = it / w.n_groups;
    x.g = it - x.ch * w.n_groups;
    x.s_lo = w.s_begin + x.ch * w.spi;
    x.s_hi = (x.s_lo + w.spi < w.s_end) ? x.s_lo + w.spi : w.s_end;
    return x;
}
// walks the attempt blocks of a CTA's items in execution order (the TMA producer runs ahead of
// the other warps across item boundaries; it needs no group state, only the stream position)
struct TcBlockWalk {
    int s, s_hi, k, ch, g, n_chunks, dch, dg;   // no division on the per-block path
    __device__ void start(const TcItems& w, int cid) {
        n_chunks = w.n_chunks();
        dch = w.n_cta / w.n_groups;
        dg = w.n_cta - dch * w.n_groups;
        ch = cid / w.n_groups;
        g = cid - ch * w.n_groups;
        k = 0;
        set_sweeps(w);
    }
    __device__ void set_sweeps(const TcItems& w) {
        s = w.s_begin + ch * w.spi;
        s_hi = (s + w.spi < w.s_end) ? s + w.spi : w.s_end;
    }
    __device__ bool valid() const { return ch < n_chunks; }
    __device__ void next(const TcItems& w, int nblk) {
        if (++k < nblk) return;
        k = 0;
        if (++s < s_hi) return;
        ch += dch;
        g += dg;
        if (g >= w.n_groups) { g -= w.n_groups; ++ch; }
        set_sweeps(w);
    }
    __device__ size_t stream_block(const TcItems& w, int nblk) const {
        return (size_t)(s - w.s_begin) * nblk + k;
    }
};
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// C = 1: one CTA holds all field columns of 16 replicas.  C = 2 / 4: a thread-block cluster holds
// 16 C replicas, CTA `crank` owns the columns [crank * n_tc/C, (crank+1) * n_tc/C) (so each SM
// streams and multiplies only 1/C of every J row for C times the replicas: 1/C of the
// shared-memory traffic per attempt); all CTAs run the same decision warps (one per 32 replicas)
// on identical inputs (raw field values are exchanged through distributed shared memory), so no
// decision has to cross the cluster.
template <int P, bool INJECT, int C>
__global__ void __launch_bounds__(tc_threads(C), 1)
sweep_tc_kernel(const SweepDev a, const __nv_bfloat16* __restrict__ Jp, const int n_tc,
                const uint16_t* __restrict__ sites_g, const int n_s, const int NS,
                const int tmem_cols, const unsigned char* __restrict__ Q,
                const unsigned char* __restrict__ tabs_g, const int s_begin, const int s_end,
                const int spi, int* __restrict__ done_g, const int dbg) {
    constexpr int LAG = 1;                   // raw field values are read LAG+1 blocks before use
    constexpr int NG = kG * C;               // replicas per group = MMA N
    constexpr int NGRP = NG / 16;            // 16-column TMEM load/store groups per tile
    constexpr uint32_t IDESC = tc::make_idesc_bf16(kTileM, NG);
    constexpr int CT = tc_chunk_tiles(C);    // tiles per ring stage
    constexpr int NDW = tc_ndw(C);           // decision warps
    constexpr int NTW = tc_ntw(C);           // dedicated threshold warps (0: the quarter warps do it)
    constexpr int kThWarps = NTW ? NTW : ((NG / 8 < 4) ? NG / 8 : 4);   // warps that draw thresholds
    auto named_sync = [](int id) { named_sync_c<C>(id); };
    constexpr int BOP = 32 * NG;             // B operand bytes per slot
    constexpr uint32_t BLBO = 16 * NG;       // B: k-group stride ((NG/8) n-groups of 128 B)
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = a.n, n_pad = a.n_pad;
    const int W = n_tc >> 5;                 // spin words per replica
    const int Wp = W + 1;                    // padded stride of a bit plane (lane = replica reads
                                             // word x of 32 planes: odd stride = 32 distinct banks)
    const int T = n_tc / kTileM;             // tiles of the whole model
    const int Tl = T / C;                    // tiles of this CTA
    // Q holds every block as (T rounded up to kChunkTiles) tiles x P planes, tile-major; this CTA
    // consumes its Tl tiles in stages of CT tiles
    const int tiles_q = (T + kChunkTiles - 1) / kChunkTiles * kChunkTiles;
    const int nchunk_l = (Tl + CT - 1) / CT;            // stages per block consumed by this CTA
    // part h = local chunks [pb(h), pb(h+1)) (some parts are empty when there are fewer chunks)
    auto pb = [nchunk_l](int h) { return (nchunk_l * h + kParts - 1) / kParts; };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = (C > 1) ? cluster_ctarank() : 0u;
    const int cols_cta = Tl * kTileM;        // field columns per CTA
    const int col0 = (int)crank * cols_cta;

    const TcSmem L = tc_layout(n_tc, P, NS, C);
    unsigned char* ring = smem + L.ring;
    unsigned char* bop_s = smem + L.bop;
    uint32_t* sbits = reinterpret_cast<uint32_t*>(smem + L.sbits);
    float* theta_s = reinterpret_cast<float*>(smem + L.theta);  // [slot][b][r]
    float* raw_s = reinterpret_cast<float*>(smem + L.raw);      // [slot][b][r]
    unsigned char* tab_s = smem + L.tab;  // [slot]{cin[16][16], ccr[16][16], dup[16]} (TMA)
    float* ztab_s = reinterpret_cast<float*>(smem + L.ztab);
    float* red = reinterpret_cast<float*>(smem + L.red);        // [rank][quarter][r]
    uint32_t* flags = reinterpret_cast<uint32_t*>(smem + L.flags);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* empty = full + kMaxStagesTc;
    uint64_t* tabbar = empty + kMaxStagesTc;   // [slot]     decision tables of a block landed (TMA)
    uint64_t* thbar = tabbar + kSlots;         // [slot]     thresholds of a block published
    uint64_t* rloc = thbar + kSlots;           // [slot][2]  this CTA's raw reads of a half done
    uint64_t* rall = rloc + kParts * kSlots;   // [slot]     all raw values of a block published
    uint64_t* decbar = rall + kSlots;          // [slot]     block decided, B operand written
    uint64_t* hdone = decbar + kSlots;         // [slot][2]  MMAs of a half of a block completed
    uint64_t* ebar = hdone + kParts * kSlots;  // energy partial sums of a sweep published
    uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + L.tptr);
    constexpr int kStageBytes = CT * P * kTileBytes;

    const int n_sweeps = a.n_sweeps;
    const int nblk = (n + kBlk - 1) / kBlk;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    // Work items: (sweep chunk ch, replica group g), numbered chunk-major; CTA (or cluster) c runs
    // items c, c + n_cta, ...  A group's state (fields, spins, energies) lives in HBM between its
    // items; done_g[g] counts the chunks of group g that are complete.  With one chunk per group
    // (spi = all sweeps, n_cta = groups) this is the plain one-CTA-per-group launch.
    const TcItems items{(a.R + NG - 1) / NG, s_begin, s_end, spi, (int)(gridDim.x / C)};
    const int cid = (int)(blockIdx.x / C);
    const int n_items = items.count();
    const int n_chunks = items.n_chunks();

    // ------------------------------------------------------------ prologue
    if (a.dbg && tid == 0 && blockIdx.x < 512) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.dbg[8192 + blockIdx.x * 4 + 0] = (long long)gt;
        a.dbg[8192 + blockIdx.x * 4 + 2] = (long long)clock64();
    }
    if (tid == 0) {
        for (int d = 0; d < NS; ++d) {
            mbar_init(&full[d], 1);
            mbar_init(&empty[d], 1);
        }
        for (int d = 0; d < kSlots; ++d) {
            mbar_init(&tabbar[d], 1);
            mbar_init(&thbar[d], kThWarps);
            for (int h = 0; h < kParts; ++h) {
                mbar_init(&rloc[kParts * d + h], 4);
                mbar_init(&hdone[kParts * d + h], 1);
            }
            mbar_init(&rall[d], 4 * kParts + (C > 1 ? 1 : 0));   // 4 warps x parts (+ the expect_tx arrival)
            mbar_init(&decbar[d], NDW);
        }
        mbar_init(ebar, 4 + (C > 1 ? 1 : 0));
        flags[0] = flags[1] = flags[2] = flags[3] = 0u;
        fence_mbar_init();
        fence_proxy_async();
    }
    for (int i = tid; i < kBlk * kBlk; i += tc_threads(C)) ztab_s[i] = 0.0f;
    if (warp == 0) {
        tc::tmem_alloc(tptr, (uint32_t)tmem_cols);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (C > 1) cluster_sync_all();   // the peers' barriers exist before anything remote arrives
    tc::fence_after_sync();
    const uint32_t tbase = *tptr;

    if (warp < 4) {
        // ======================================================== QUARTER WARPS
        const int q = warp;
        const uint32_t tq = tbase + ((uint32_t)(q * 32) << 16);
        int kg = 0;
#pragma unroll 1
        for (int it = cid; it < n_items; it += items.n_cta) {
        const TcItem item = tc_item(items, it);
        const int rep0 = item.g * NG;
        const int g_act = min(NG, a.R - rep0);
        // ---- item prologue: the group's previous chunk is complete, then state HBM -> SM
        SG_ISTAMP(0);
        if (item.ch > 0 && tid == 0) {   // every CTA of the group's previous chunk has published
            while (ld_acquire_gpu(done_g + item.g) < C * item.ch) __nanosleep(256);
        }
        named_sync(2);
        SG_ISTAMP(1);
        // spin bit planes from int8 spins (every CTA of a group keeps all NG planes); four words
        // (eight 16-byte loads) in flight per thread
        for (int w0 = tid; w0 < NG * W; w0 += 4 * 128) {
            uint4 lo[4], hi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int w = w0 + u * 128;
                const int r = w / W, word = w - r * W;
                lo[u] = hi[u] = make_uint4(0u, 0u, 0u, 0u);
                if (w < NG * W && r < g_act && word * 32 < n_pad) {   // n_tc may exceed n_pad
                    const uint4* src =
                        reinterpret_cast<const uint4*>(a.spins + (size_t)(rep0 + r) * n_pad + word * 32);
                    lo[u] = __ldcg(src);
                    hi[u] = __ldcg(src + 1);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int w = w0 + u * 128;
                const int r = w / W, word = w - r * W;
                if (w >= NG * W) break;
                uint32_t bits = 0xFFFFFFFFu;
                if (r < g_act && word * 32 < n_pad) {
                    const uint32_t x[8] = {lo[u].x, lo[u].y, lo[u].z, lo[u].w, hi[u].x, hi[u].y, hi[u].z, hi[u].w};
                    bits = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t up = (~x[j]) & 0x80808080u;  // byte >= 0  <=> spin up
                        const uint32_t nib =
                            ((up >> 7) & 1u) | ((up >> 14) & 2u) | ((up >> 21) & 4u) | ((up >> 28) & 8u);
                        bits |= nib << (4 * j);
                    }
                }
                sbits[r * Wp + word] = bits;
            }
        }
        SG_ISTAMP(2);
        // resident fields -> TMEM, 32+ loads per thread in flight
        constexpr int GS = (NGRP < 4) ? NGRP : 4;      // 16-replica groups per pass (<= 64 registers)
        constexpr int TU = (GS >= 4) ? 1 : 2;
        for (int t = 0; t < Tl; t += TU)
#pragma unroll 1
        for (int g0 = 0; g0 < NGRP; g0 += GS) {
            float v[TU][GS][16];
#pragma unroll
            for (int u = 0; u < TU; ++u) {
                const int col = col0 + (t + u) * kTileM + q * 32 + lane;
#pragma unroll
                for (int gi = 0; gi < GS; ++gi)
#pragma unroll
                    for (int r = 0; r < 16; ++r)
                        v[u][gi][r] = (t + u < Tl && (g0 + gi) * 16 + r < g_act && col < n_pad)
                                          ? __ldcg(a.fields + (size_t)(rep0 + (g0 + gi) * 16 + r) * n_pad + col) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < TU; ++u)
                if (t + u < Tl) {
#pragma unroll
                    for (int gi = 0; gi < GS; ++gi) tc::tmem_st16(tq + (t + u) * NG + (g0 + gi) * 16, v[u][gi]);
                }
        }
        tc::wait_st();
        tc::fence_before_sync();
        named_sync(2);   // bit planes complete (decision warp reads them)
        tc::fence_after_sync();
        SG_ISTAMP(3);
        // thresholds of block (s, kb) -> theta_s[slot]: thread (qq, r) covers attempts 4qq..4qq+3.
        // They depend on nothing but the counters, so they are drawn ONE BLOCK AHEAD, after the raw
        // reads of the current block (which sit on the critical path between two blocks' MMAs).
        auto draw_thresholds = [&](int s_t, int kb_t, int slot_t) {
            if (NTW) return;   // drawn by the threshold warps
            if (!INJECT) {
                const unsigned long long sa_t = a.sweep_base + (unsigned long long)s_t;
                const int i0_t = kb_t * kBlk;
#pragma unroll 1
                for (int tt = tid; tt < 4 * NG; tt += 128) {
                    const int qq = tt / NG, r = tt - qq * NG;
                    const int ia = i0_t + qq * 4;
                    if (r < g_act && ia < n) {
                        const int rep = rep0 + r;
                        const float Tm = (float)a.temps[(long long)s_t * a.t_ss + (long long)rep * a.t_rs];
                        const uint4 x = philox4x32_10(
                            make_uint4((uint32_t)(rep + a.rep_base), (uint32_t)sa_t, (uint32_t)(sa_t >> 32),
                                       (uint32_t)(ia >> 2)), key);
                        const uint32_t vv[4] = {x.x, x.y, x.z, x.w};
                        float* dst = theta_s + slot_t * kBlk * NG + (qq * 4) * NG + r;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float u = u01(vv[e]);
                            float th;
                            if (a.rule == 0) th = -__logf(u) * Tm;
                            else th = 0.5f * Tm * (__logf(u) - __logf(1.0f - u));
                            dst[e * NG] = th;
                        }
                    }
                }
            }
            __syncwarp();
            if (warp < kThWarps && lane == 0) mbar_arrive(&thbar[slot_t]);
        };
        draw_thresholds(item.s_lo, 0, kg & (kSlots - 1));
        // the 16 sites of a block come from global memory (L2: ~500 clocks): they are fetched one
        // block ahead, like the thresholds
        uint4 nsq0, nsq1;
        int n_my_site;
        auto fetch_sites = [&](int s_t, int kb_t) {
            const uint16_t* st_t = sites_g + (size_t)s_t * n_s + kb_t * kBlk;
            nsq0 = *reinterpret_cast<const uint4*>(st_t);
            nsq1 = *reinterpret_cast<const uint4*>(st_t + 8);
            n_my_site = (int)st_t[lane & 15];
        };
        fetch_sites(item.s_lo, 0);
#pragma unroll 1
        for (int s = item.s_lo; s < item.s_hi; ++s) {
#pragma unroll 1
            for (int kb = 0; kb < nblk; ++kb, ++kg) {
                const int slot = kg & (kSlots - 1);
                const int i0 = kb * kBlk;
                const int nbk = min(kBlk, n - i0);
                if (warp == 0) SG_STAMP(0);
                const uint4 sq0 = nsq0, sq1 = nsq1;
                const int my_site = n_my_site;
                const uint32_t swq[8] = {sq0.x, sq0.y, sq0.z, sq0.w, sq1.x, sq1.y, sq1.z, sq1.w};
                {   // next block of this item (its data are consumed in the next iteration)
                    const bool more = kb + 1 < nblk;
                    if (more || s + 1 < item.s_hi) fetch_sites(more ? s : s + 1, more ? kb + 1 : 0);
                }
                if (C > 1 && warp == 0 && lane == 0) {
                    // the peers will store the raw values of the sites they own straight into this
                    // CTA's raw_s[slot], counting bytes on rall[slot]
                    int n_remote = 0;
#pragma unroll
                    for (int b = 0; b < kBlk; ++b) {
                        const int site = (int)((swq[b >> 1] >> (16 * (b & 1))) & 0xFFFFu);
                        const int lcr = site - col0;
                        n_remote += (b < nbk && !(lcr >= 0 && lcr < cols_cta)) ? 1 : 0;
                    }
                    mbar_arrive_expect_tx(&rall[slot], (uint32_t)(n_remote * NG * 4));
                }
                if (warp == 0) SG_STAMP(1);
                // --- raw field values of the block's sites in this CTA's columns, as of the end of
                // block kg-LAG-1: the sites of a column half are read as soon as that half of that
                // block has landed (the other half may still be executing), and the same half of
                // block kg-LAG is not issued before the read is done.  (LAG = 2 in the cluster
                // variant gives the cross-CTA exchange and the decision a whole block of slack.)
                float* rawb = raw_s + slot * kBlk * NG;
                uint32_t raw_peer[C > 1 ? C - 1 : 1], bar_peer[C > 1 ? C - 1 : 1];
                if (NGRP >= 4) {
#pragma unroll
                    for (int pp = 1; pp < C; ++pp) {
                        const uint32_t peer = (crank + (uint32_t)pp) & (uint32_t)(C - 1);
                        raw_peer[pp - 1] = map_to_rank(rawb, peer);
                        bar_peer[pp - 1] = map_to_rank(&rall[slot], peer);
                    }
                }
                // lane b keeps site b; the loops below stay rolled (this is hot code shared with six
                // other warps' loops in one instruction cache: unrolled 16 x 2 it made the cluster
                // variant's working set spill out of it)
                // lane b (< 16) owns site b of the block and decides, before any waiting, whether
                // this warp reads it (its column is in this CTA and in this warp's TMEM lane
                // quarter) and in which column half: two ballots replace a 16-step scan per half
                // in what is, in the cluster variant, the critical path between two blocks' MMAs
                const int my_lc = my_site - col0;          // column inside this CTA
                const bool my_take = lane < nbk && ((C == 1) || (my_lc >= 0 && my_lc < cols_cta)) &&
                                     ((my_lc >> 5) & 3) == q && !(dbg & 4);
                int my_part = 0;
                {
                    const int my_chunk = my_lc / (CT * kTileM);
#pragma unroll
                    for (int h = 1; h < kParts; ++h) my_part = (my_chunk >= pb(h)) ? h : my_part;
                }
                // (rolled: this is hot code shared with the other warps' loops in one instruction cache)
#pragma unroll 1
                for (int h = 0; h < kParts; ++h) {
                    uint32_t todo = __ballot_sync(0xFFFFFFFFu, my_take && my_part == h);
                    if (kg >= LAG + 1)
                        mbar_wait(&hdone[kParts * ((kg - LAG - 1) & (kSlots - 1)) + h],
                                  (uint32_t)((kg - LAG - 1) >> 2) & 1u);
                    tc::fence_after_sync();
                    if (warp == 0 && (h == 0 || h == kParts - 1)) SG_STAMP(h == 0 ? 2 : 3);

#pragma unroll 1
                    while (todo) {
                        const int b = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const int lc = __shfl_sync(0xFFFFFFFFu, my_lc, b);
                        float* dl = rawb + b * NG;
                        if (NGRP >= 4) {
                            // 64 / 128 replicas: four loads in flight per wait; the owning lane puts
                            // the row into this CTA's table, then EVERY lane forwards its 8 (16)
                            // bytes of it to each peer (asynchronous remote stores that count their
                            // bytes on the peer's rall[slot]) -- C-1 stores per lane instead of
                            // 16 (C-1) from one lane, in the critical path between two blocks' MMAs
#pragma unroll 1
                            for (int g0 = 0; g0 < NGRP; g0 += 4) {
                                float v[4][16];
#pragma unroll
                                for (int gi = 0; gi < 4; ++gi) tc::tmem_ld16(tq + (lc >> 7) * NG + (g0 + gi) * 16, v[gi]);
                                tc::wait_ld();
                                if (lane == (lc & 31)) {
#pragma unroll
                                    for (int gi = 0; gi < 4; ++gi) {
                                        float4* d4 = reinterpret_cast<float4*>(dl + (g0 + gi) * 16);
                                        d4[0] = make_float4(v[gi][0], v[gi][1], v[gi][2], v[gi][3]);
                                        d4[1] = make_float4(v[gi][4], v[gi][5], v[gi][6], v[gi][7]);
                                        d4[2] = make_float4(v[gi][8], v[gi][9], v[gi][10], v[gi][11]);
                                        d4[3] = make_float4(v[gi][12], v[gi][13], v[gi][14], v[gi][15]);
                                    }
                                }
                            }
                            __syncwarp();
                            constexpr int FW = NG / 32;   // floats of the row each lane forwards
                            const uint32_t off = (uint32_t)((b * NG + FW * lane) * 4);
                            if (FW == 2) {
                                const float2 mine = *reinterpret_cast<const float2*>(dl + 2 * lane);
#pragma unroll
                                for (int pp = 1; pp < C; ++pp) st_async_f2(raw_peer[pp - 1] + off, mine, bar_peer[pp - 1]);
                            } else {
                                const float4 mine = *reinterpret_cast<const float4*>(dl + 4 * lane);
#pragma unroll
                                for (int pp = 1; pp < C; ++pp) st_async_f4(raw_peer[pp - 1] + off, mine, bar_peer[pp - 1]);
                            }
                        } else
#pragma unroll
                        for (int gi = 0; gi < NGRP; ++gi) {
                            float v[16];
                            tc::tmem_ld16(tq + (lc >> 7) * NG + gi * 16, v);
                            tc::wait_ld();
                            if (lane == (lc & 31)) {
                                float4* d4 = reinterpret_cast<float4*>(dl + gi * 16);
                                d4[0] = make_float4(v[0], v[1], v[2], v[3]);
                                d4[1] = make_float4(v[4], v[5], v[6], v[7]);
                                d4[2] = make_float4(v[8], v[9], v[10], v[11]);
                                d4[3] = make_float4(v[12], v[13], v[14], v[15]);
                                if (C > 1) {
                                    // the same 64 bytes straight from the registers into every
                                    // peer's shared memory (asynchronous remote stores that count
                                    // their bytes on the peer's rall[slot])
#pragma unroll
                                    for (int pp = 1; pp < C; ++pp) {
                                        const uint32_t peer = (crank + (uint32_t)pp) & (uint32_t)(C - 1);
                                        const uint32_t ra = map_to_rank(dl + gi * 16, peer);
                                        const uint32_t rb = map_to_rank(&rall[slot], peer);
                                        st_async_f4(ra, make_float4(v[0], v[1], v[2], v[3]), rb);
                                        st_async_f4(ra + 16, make_float4(v[4], v[5], v[6], v[7]), rb);
                                        st_async_f4(ra + 32, make_float4(v[8], v[9], v[10], v[11]), rb);
                                        st_async_f4(ra + 48, make_float4(v[12], v[13], v[14], v[15]), rb);
                                    }
                                }
                            }
                        }
                    }
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&rloc[kParts * slot + h]);
                        mbar_arrive(&rall[slot]);
                    }
                }
                if (warp == 0) SG_STAMP(15);
                // thresholds of the next block of this item
                if (kb + 1 < nblk) draw_thresholds(s, kb + 1, (kg + 1) & (kSlots - 1));
                else if (s + 1 < item.s_hi) draw_thresholds(s + 1, 0, (kg + 1) & (kSlots - 1));
            }
            // ---- end of sweep: energy partial sums over this CTA's columns
            mbar_wait(&hdone[kParts * ((kg - 1) & (kSlots - 1)) + kParts - 1], (uint32_t)((kg - 1) >> 2) & 1u);
            tc::fence_after_sync();
            named_sync(1);  // decision warp has flipped the last spins of the sweep
            {
                constexpr int PG = (NGRP < 4) ? NGRP : 4;   // 16-replica groups per pass
                if (C > 1 && q == 0 && lane == 0) mbar_arrive_expect_tx(ebar, (uint32_t)((C - 1) * 4 * NG * 4));
#pragma unroll 1
                for (int g0 = 0; g0 < NGRP; g0 += PG) {
                    float part[PG * 16];
#pragma unroll
                    for (int r = 0; r < PG * 16; ++r) part[r] = 0.0f;
                    for (int t = 0; t < Tl; ++t) {
                        const int colb = col0 + t * kTileM + q * 32;   // first column of this lane group
                        const float hv = (colb + lane < n_pad) ? a.h[colb + lane] : 0.0f;
#pragma unroll
                        for (int gi = 0; gi < PG; ++gi) {
                            float f[16];
                            tc::tmem_ld16(tq + t * NG + (g0 + gi) * 16, f);
                            tc::wait_ld();
#pragma unroll
                            for (int r = 0; r < 16; ++r) {
                                const uint32_t w = sbits[((g0 + gi) * 16 + r) * Wp + (colb >> 5)];
                                const float tt = f[r] + hv;
                                part[gi * 16 + r] += ((w >> lane) & 1u) ? tt : -tt;
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < PG * 16; ++r) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) part[r] += __shfl_xor_sync(0xFFFFFFFFu, part[r], o);
                    }
                    if (lane == 0) {
                        float* dst = red + ((int)crank * 4 + q) * NG + g0 * 16;
#pragma unroll
                        for (int r = 0; r < PG * 16; ++r) dst[r] = part[r];
                        if (C > 1) {
#pragma unroll 1
                            for (int pp = 1; pp < C; ++pp) {
                                const uint32_t peer = (crank + (uint32_t)pp) & (uint32_t)(C - 1);
                                const uint32_t ra = map_to_rank(dst, peer);
                                const uint32_t rb = map_to_rank(ebar, peer);
#pragma unroll
                                for (int r4 = 0; r4 < PG * 4; ++r4)
                                    st_async_f4(ra + 16 * r4, make_float4(part[4 * r4], part[4 * r4 + 1],
                                                                         part[4 * r4 + 2], part[4 * r4 + 3]), rb);
                            }
                        }
                    }
                }
                if (lane == 0) mbar_arrive(ebar);
            }
            tc::fence_before_sync();
            named_sync(3);  // flags[dw] = mask of the replicas of decision warp dw that improved
            if ((flags[0] | flags[1] | flags[2] | flags[3]) != 0u && crank == 0) {
                for (int w = tid; w < NG * W; w += 128) {
                    const int r = w / W, word = w - r * W;
                    if (((flags[r >> 5] >> (r & 31)) & 1u) && word * 32 < n_pad)
                        store_spin_word(a.best_spins + (size_t)(rep0 + r) * n_pad + word * 32, sbits[r * Wp + word]);
                }
            }
            named_sync(1);  // bit planes may be modified again
        }
        // ---- item epilogue: fields (this CTA's columns) and spins back to HBM
        SG_ISTAMP(4);
        for (int t = 0; t < Tl; ++t) {
            const int col = col0 + t * kTileM + q * 32 + lane;
#pragma unroll
            for (int gi = 0; gi < NGRP; ++gi) {
                float v[16];
                tc::tmem_ld16(tq + t * NG + gi * 16, v);
                tc::wait_ld();
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    if (gi * 16 + r < g_act && col < n_pad)
                        __stcg(a.fields + (size_t)(rep0 + gi * 16 + r) * n_pad + col, v[r]);
            }
        }
        if (crank == 0) {
            for (int w = tid; w < NG * W; w += 128) {
                const int r = w / W, word = w - r * W;
                if (r < g_act && word * 32 < n_pad)
                    store_spin_word(a.spins + (size_t)(rep0 + r) * n_pad + word * 32, sbits[r * Wp + word]);
            }
        }
        SG_ISTAMP(5);
        __threadfence();
        SG_ISTAMP(6);
        named_sync(2);   // every store of the item (quarter warps and decision warp) is fenced
        SG_ISTAMP(7);
        if (tid == 0 && n_chunks > 1) red_release_gpu_add(done_g + item.g, 1);
        }  // items
    } else if (warp == (NTW ? 6 + NDW + 1 : 4)) {
        // ======================================================== PRODUCER (TMA)
        if (lane == 0) {
            int stage = 0;
            uint32_t epar = 1;  // a fresh barrier passes a wait on the "previous" phase
            int kg = 0;
            TcBlockWalk wt, wq;   // table prefetch position (LAG+1 blocks ahead) and stream position
            wt.start(items, cid);
            wq.start(items, cid);
            int jt = 0;
            // decision tables of the CTA's block j -> slot j % 4 (the slot's previous user, block
            // j-4, has been decided long before the operand stream reaches block j-2)
            auto issue_tables = [&]() {
                if (!wt.valid()) return;
                const int sl = jt & (kSlots - 1);
                if (jt >= kSlots) mbar_wait(&decbar[sl], (uint32_t)((jt - kSlots) >> 2) & 1u);
                mbar_arrive_expect_tx(&tabbar[sl], (uint32_t)kTabBytes);
                bulk_g2s(tab_s + sl * kTabBytes, tabs_g + wt.stream_block(items, nblk) * kTabBytes,
                         (uint32_t)kTabBytes, &tabbar[sl]);
                wt.next(items, nblk);
                ++jt;
            };
#pragma unroll
            for (int j = 0; j <= LAG; ++j) issue_tables();
#pragma unroll 1
            for (; wq.valid(); wq.next(items, nblk), ++kg) {
                issue_tables();
                const unsigned char* src =
                    Q + (wq.stream_block(items, nblk) * tiles_q + (size_t)crank * Tl) * (size_t)(P * kTileBytes);
#pragma unroll 1
                for (int c = 0; c < nchunk_l; ++c) {
                    mbar_wait(&empty[stage], epar);
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)kStageBytes);
                    bulk_g2s(ring + (size_t)stage * kStageBytes, (dbg & 1) ? Q : src + (size_t)c * kStageBytes,
                             (uint32_t)kStageBytes, &full[stage]);
                    if (++stage == NS) { stage = 0; epar ^= 1u; }
                }
            }
        }
    } else if (NTW && (warp == 4 || warp == 6 + NDW)) {
        // ======================================================== THRESHOLD WARPS
        // thresholds of block kg -> theta_s[kg % 4]: unit (qq, r) covers attempts 4qq..4qq+3 of
        // replica r (one Philox call).  They depend on nothing but the counters; the only
        // constraint is the slot's previous user, block kg-4, having been decided.
        const int tw = (warp == 4) ? 0 : 1;
        TcBlockWalk wk;
        wk.start(items, cid);
        int kg = 0;
#pragma unroll 1
        for (; wk.valid(); wk.next(items, nblk), ++kg) {
            const int slot = kg & (kSlots - 1);
            if (kg >= kSlots) mbar_wait(&decbar[slot], (uint32_t)((kg - kSlots) >> 2) & 1u);
            if (!INJECT) {
                const int rep0 = wk.g * NG;
                const int g_act = min(NG, a.R - rep0);
                const unsigned long long sa_t = a.sweep_base + (unsigned long long)wk.s;
                const int i0_t = wk.k * kBlk;
#pragma unroll 1
                for (int tt = tw * 32 + lane; tt < 4 * NG; tt += 32 * (NTW ? NTW : 1)) {
                    const int qq = tt / NG, r = tt - qq * NG;
                    const int ia = i0_t + qq * 4;
                    if (r < g_act && ia < n) {
                        const int rep = rep0 + r;
                        const float Tm = (float)a.temps[(long long)wk.s * a.t_ss + (long long)rep * a.t_rs];
                        const uint4 x = philox4x32_10(
                            make_uint4((uint32_t)(rep + a.rep_base), (uint32_t)sa_t, (uint32_t)(sa_t >> 32),
                                       (uint32_t)(ia >> 2)), key);
                        const uint32_t vv[4] = {x.x, x.y, x.z, x.w};
                        float* dst = theta_s + slot * kBlk * NG + (qq * 4) * NG + r;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float u = u01(vv[e]);
                            float th;
                            if (a.rule == 0) th = -__logf(u) * Tm;
                            else th = 0.5f * Tm * (__logf(u) - __logf(1.0f - u));
                            dst[e * NG] = th;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&thbar[slot]);
        }
    } else if (warp == 5 || (warp >= 7 && warp < 6 + NDW)) {
        // ======================================================== DECISION WARPS
        // (one per 32 replicas of the group: warp 5, and warps 7.. for the replicas beyond 32)
        const int dw = (warp == 5) ? 0 : warp - 6;
        const int rl = dw * 32 + lane;           // replica of the group this lane decides
        const int r = rl & (NG - 1);
        static_assert(LAG == 1, "the decision warps carry one block of corrections");
        // cn[b] = correction of site b of the NEXT block for the flips of the current block (they
        // are not in the raw values the next block reads): accumulated while the current block's
        // attempts are decided -- row a of the next block's table as soon as attempt a is known,
        // independent work that fills the latency gaps of the sequential decision chain (it used
        // to be a separate 16 x 16 pass at the start of every block, ~900 clocks of this warp's
        // serial budget).  Same FMAs in the same order as before.
        float2 cn[kBlk / 2];
        int kg = 0;
        uint32_t epar = 0;
#pragma unroll 1
        for (int it = cid; it < n_items; it += items.n_cta) {
        const TcItem item = tc_item(items, it);
        const int rep0 = item.g * NG;
        const int g_act = min(NG, a.R - rep0);
        const bool active = rl < g_act;
        named_sync(2);   // the group's previous chunk is complete
        float best_e = 3.0e38f, cur_e = 0.0f;
        unsigned int n_acc = 0;
        if (active) {
            cur_e = __ldcg(a.energy + rep0 + rl);
            best_e = a.track_best ? __ldcg(a.best_energy + rep0 + rl) : 3.0e38f;
        }
        named_sync(2);   // bit planes loaded
#pragma unroll
        for (int b = 0; b < kBlk / 2; ++b) cn[b] = make_float2(0.0f, 0.0f);
        // the block's 16 sites come from global memory (L2, ~500 clocks): fetched one block ahead
        uint4 ns0, ns1;
        auto fetch_sites = [&](int s_t, int k_t) {
            const uint16_t* st_t = sites_g + (size_t)s_t * n_s + k_t * kBlk;
            ns0 = *reinterpret_cast<const uint4*>(st_t);
            ns1 = *reinterpret_cast<const uint4*>(st_t + 8);
        };
        fetch_sites(item.s_lo, 0);
#pragma unroll 1
        for (int s = item.s_lo; s < item.s_hi; ++s) {
            double dT = 1.0;
            if (INJECT && active)
                dT = a.temps[(long long)s * a.t_ss + (long long)(rep0 + rl) * a.t_rs];
#pragma unroll 1
            for (int k = 0; k < nblk; ++k, ++kg) {
                const int slot = kg & (kSlots - 1);
                const uint32_t par = (uint32_t)(kg >> 2) & 1u;
                const int i0 = k * kBlk;
                const int nbk = min(kBlk, n - i0);
                // the 16 sites of the block (uniform)
                const uint4 s0 = ns0, s1 = ns1;
                const uint32_t sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                {
                    const bool more = k + 1 < nblk;
                    if (more || s + 1 < item.s_hi) fetch_sites(more ? s : s + 1, more ? k + 1 : 0);
                }
                int site[kBlk];
#pragma unroll
                for (int b = 0; b < kBlk; ++b) site[b] = (int)((sw[b >> 1] >> (16 * (b & 1))) & 0xFFFFu);
                float uu[kBlk];
                if (INJECT) {
#pragma unroll
                    for (int b = 0; b < kBlk; ++b)
                        uu[b] = (active && b < nbk)
                                    ? a.uniforms[((size_t)(rep0 + rl) * n_sweeps + s) * n + i0 + b]
                                    : 0.0f;
                }
                SG_STAMP(4);
                // (these three barrier tests on three lanes instead of one after the other: slower, 19.3
                // instead of 20.1 G attempts/s -- the thresholds are not always there yet, and a lone
                // polling lane notices late)
                if (k == 0) mbar_wait(&tabbar[slot], par);   // (k > 0: waited for as the previous block's "next" table)
                mbar_wait(&thbar[slot], par);
                const float* rawp = raw_s + slot * kBlk * NG + r;
                const float* thp = theta_s + slot * kBlk * NG + r;
                const float4* cin4 = reinterpret_cast<const float4*>(tab_s + slot * kTabBytes);
                const uint32_t* dup_p = reinterpret_cast<const uint32_t*>(tab_s + slot * kTabBytes) + 3 * kBlk * kBlk;
                // the next block's cross table (couplings from this block's sites into the next
                // block's; none across a sweep boundary: the first block of a sweep reads its raw
                // values after everything before it has landed)
                const bool has_next = k + 1 < nblk;
                if (has_next) mbar_wait(&tabbar[(kg + 1) & (kSlots - 1)], (uint32_t)((kg + 1) >> 2) & 1u);
                // (no next block: a table of zeros, so that the loads below are unconditional and
                // can be scheduled across the whole unrolled attempt chain)
                const float4* nx4 = has_next
                    ? reinterpret_cast<const float4*>(tab_s + ((kg + 1) & (kSlots - 1)) * kTabBytes) + (kBlk * kBlk / 4)
                    : reinterpret_cast<const float4*>(ztab_s);
                // (pairs of fields per register pair: one FFMA2 = two independent fp32 FMAs, same
                // results as scalar fmaf, half the instructions of this warp's serial budget)
                float2 v2[kBlk / 2];
#pragma unroll
                for (int b = 0; b < kBlk / 2; ++b) {
                    v2[b] = cn[b];
                    cn[b] = make_float2(0.0f, 0.0f);
                }
                float th[kBlk];
                uint32_t w0[kBlk], dup[kBlk];
#pragma unroll
                for (int b = 0; b < kBlk; ++b) {
                    th[b] = INJECT ? 0.0f : thp[b * NG];
                    w0[b] = sbits[r * Wp + (site[b] >> 5)];
                    dup[b] = dup_p[b];
                }
                SG_STAMP(5);
                mbar_wait(&rall[slot], par);
                SG_STAMP(6);
#pragma unroll
                for (int b = 0; b < kBlk / 2; ++b) {
                    v2[b].x += rawp[(2 * b) * NG];
                    v2[b].y += rawp[(2 * b + 1) * NG];
                }
                // the 16 attempts of the block, strictly in order, registers only
                uint32_t myflips = 0;
                float d[kBlk];
                uint32_t anydup = 0u;
#pragma unroll
                for (int b = 0; b < kBlk; ++b) anydup |= dup[b];
                if (!INJECT && a.rule == 0 && anydup == 0u) {
                    // Fast path (Metropolis, production RNG, no site twice in the block -- 97 % of
                    // the blocks at N = 4096): the spin of every attempt is known up front and
                    // inactive attempts get a threshold of -inf, so the chain from one decision to
                    // the next is FFMA2 -> FMUL -> FSETP -> FSEL instead of running through the
                    // flip mask, a population count and three predicate combinations as well.
                    float sg2[kBlk], the[kBlk];
#pragma unroll
                    for (int aa = 0; aa < kBlk; ++aa) {
                        sg2[aa] = ((w0[aa] >> (site[aa] & 31)) & 1u) ? 2.0f : -2.0f;   // 2 s
                        the[aa] = (active && aa < nbk) ? th[aa] : -INFINITY;
                    }
#pragma unroll
                    for (int aa = 0; aa < kBlk; ++aa) {
                        const float fv = (aa & 1) ? v2[aa >> 1].y : v2[aa >> 1].x;
                        const bool flip = sg2[aa] * fv < the[aa];       // dE = 2 s f < -T ln u
                        const float da = flip ? -sg2[aa] : 0.0f;
                        d[aa] = da;
                        myflips |= flip ? (1u << aa) : 0u;
                        const float2 da2 = make_float2(da, da);
#pragma unroll
                        for (int p2 = (aa + 1) / 2; p2 < kBlk / 2; ++p2) {
                            const float2 c2 = *reinterpret_cast<const float2*>(
                                reinterpret_cast<const float*>(cin4) + aa * kBlk + 2 * p2);
                            v2[p2] = __ffma2_rn(da2, c2, v2[p2]);
                        }
#pragma unroll
                        for (int b4 = 0; b4 < kBlk / 4; ++b4) {
                            const float4 c4 = nx4[aa * 4 + b4];
                            cn[2 * b4 + 0] = __ffma2_rn(da2, make_float2(c4.x, c4.y), cn[2 * b4 + 0]);
                            cn[2 * b4 + 1] = __ffma2_rn(da2, make_float2(c4.z, c4.w), cn[2 * b4 + 1]);
                        }
                    }
                } else {
#pragma unroll
                for (int aa = 0; aa < kBlk; ++aa) {
                    const bool up = (((w0[aa] >> (site[aa] & 31)) ^ (uint32_t)__popc(myflips & dup[aa])) & 1u) != 0u;
                    bool flip = false;
                    const float fv = (aa & 1) ? v2[aa >> 1].y : v2[aa >> 1].x;
                    if (!INJECT) {
                        if (a.rule == 0) {
                            const float x = up ? 2.0f * fv : -2.0f * fv;  // dE = 2 s f
                            flip = x < th[aa];
                        } else {
                            flip = ((fv > th[aa]) != up);
                        }
                    } else {
                        if (a.rule == 0) {
                            const float x = up ? 2.0f * fv : -2.0f * fv;
                            flip = (x <= 0.0f) || (uu[aa] < expf((float)(-(double)x / dT)));
                        } else {
                            const float arg = (a.rule == 1) ? (float)(-2.0 * (double)fv / dT)
                                                            : (float)(-2.0 * (1.0 / dT) * (double)fv);
                            const float p_up = 1.0f / (1.0f + expf(arg));
                            flip = ((uu[aa] < p_up) != up);
                        }
                    }
                    flip = flip && active && (aa < nbk);
                    const float da = flip ? (up ? -2.0f : 2.0f) : 0.0f;
                    d[aa] = da;
                    myflips |= flip ? (1u << aa) : 0u;
                    // bring the later sites of the block up to date (row aa of the in-block table)
#pragma unroll
                    // (values of sites already decided are dead, so whole pairs are updated)
                    const float2 da2 = make_float2(da, da);
#pragma unroll
                    for (int p2 = (aa + 1) / 2; p2 < kBlk / 2; ++p2) {
                        const float2 c2 = *reinterpret_cast<const float2*>(
                            reinterpret_cast<const float*>(cin4) + aa * kBlk + 2 * p2);
                        v2[p2] = __ffma2_rn(da2, c2, v2[p2]);
                    }
#pragma unroll
                    for (int b4 = 0; b4 < kBlk / 4; ++b4) {
                        const float4 c4 = nx4[aa * 4 + b4];
                        cn[2 * b4 + 0] = __ffma2_rn(da2, make_float2(c4.x, c4.y), cn[2 * b4 + 0]);
                        cn[2 * b4 + 1] = __ffma2_rn(da2, make_float2(c4.z, c4.w), cn[2 * b4 + 1]);
                    }
                }
                }
                n_acc += (unsigned int)__popc(myflips);
                SG_STAMP(7);
                // B operand (K-major bf16): byte(n, k) = (n/8)*128 + (k/8)*BLBO + (n%8)*16 + (k%8)*2
                if (rl < NG) {
                    unsigned char* bo = bop_s + slot * BOP + (rl & 7) * 16 + (rl >> 3) * kBSbo;
                    *reinterpret_cast<uint4*>(bo) =
                        make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]),
                                   pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
                    *reinterpret_cast<uint4*>(bo + BLBO) =
                        make_uint4(pack_bf16x2(d[8], d[9]), pack_bf16x2(d[10], d[11]),
                                   pack_bf16x2(d[12], d[13]), pack_bf16x2(d[14], d[15]));
                }
                fence_proxy_async();  // B operand visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&decbar[slot]);
                // spin bit planes (lane r owns plane r; XOR commutes, so order is irrelevant)
                if (rl < NG) {
#pragma unroll
                    for (int aa = 0; aa < kBlk; ++aa)
                        atomicXor(&sbits[r * Wp + (site[aa] >> 5)],
                                  ((myflips >> aa) & 1u) << (site[aa] & 31));
                }
                SG_STAMP(8);
            }
            // ---- end of sweep: energy, best tracking
            named_sync(1);
            mbar_wait(ebar, epar);
            epar ^= 1u;
            bool improved = false;
            if (active) {
                float acc = 0.0f;
#pragma unroll
                for (int rk = 0; rk < C; ++rk) {
                    const float* pr = red + rk * 4 * NG + r;
                    acc += (pr[0] + pr[NG]) + (pr[2 * NG] + pr[3 * NG]);
                }
                cur_e = -0.5f * acc;
                if (a.energy_trace && crank == 0) a.energy_trace[(size_t)s * (a.trace_ld ? a.trace_ld : a.R) + rep0 + rl] = cur_e;
                if (a.track_best && cur_e < best_e) {
                    best_e = cur_e;
                    improved = true;
                }
            }
            const uint32_t im = __ballot_sync(0xFFFFFFFFu, improved);
            if (lane == 0) flags[dw] = im;
            named_sync(3);
            named_sync(1);
        }
        if (active && crank == 0) {
            __stcg(a.energy + rep0 + rl, cur_e);
            if (a.track_best) __stcg(a.best_energy + rep0 + rl, best_e);
            __stcg(a.accepted + rep0 + rl, __ldcg(a.accepted + rep0 + rl) + (unsigned long long)n_acc);
        }
        __threadfence();
        named_sync(2);   // item complete (tid 0 publishes the group's progress)
        }  // items
    } else {
        // ======================================================== MMA ISSUER
        // The whole warp runs the loops (descriptor arithmetic stays warp-uniform); one elected
        // lane issues the tcgen05 instructions.
        int stage = 0;
        uint32_t fpar = 0;
        int kg = 0;
        int n_own = 0;   // sweeps this CTA runs, over all of its items
        for (int it = cid; it < n_items; it += items.n_cta) {
            const TcItem item = tc_item(items, it);
            n_own += item.s_hi - item.s_lo;
        }
#pragma unroll 1
        for (int s = 0; s < n_own; ++s) {
#pragma unroll 1
            for (int k = 0; k < nblk; ++k, ++kg) {
                const int slot = kg & (kSlots - 1);
                SG_STAMP(9);
                mbar_wait(&decbar[slot], (uint32_t)(kg >> 2) & 1u);
                SG_STAMP(10);
                const uint64_t bdesc = tc::make_smem_desc(smem_u32(bop_s + slot * BOP), BLBO, kBSbo);
#pragma unroll 1
                for (int h = 0; h < kParts; ++h) {
                    // this CTA's raw reads of block k+LAG in this column part must be done before
                    // this block's update of the part is issued
                    if (h == 1) SG_STAMP(13);
                    if (k + LAG < nblk)
                        mbar_wait(&rloc[kParts * ((kg + LAG) & (kSlots - 1)) + h],
                                  (uint32_t)((kg + LAG) >> 2) & 1u);
                    if (h == 0) SG_STAMP(11);
                    if (h == 1) SG_STAMP(14);
                    const int c_begin = pb(h), c_end = pb(h + 1);
                    // the ring stages of this part: lane l waits for stage l (the barrier tests of all
                    // stages run side by side instead of one ~65-clock round trip after the other),
                    // then ONE elected lane issues the whole part: MMAs, the stage releases and the
                    // part's commit
                    // (only when the ring holds more than two parts: with a shallow ring -- pairs and
                    // single CTAs at N = 4096 -- waiting for a whole part before its first MMA would
                    // take the overlap of copies and MMAs away; those issue stage by stage)
                    // (compiled for the cluster forms only: the other instantiations keep their loop small.
                    // Also tried: the decision and read barriers on lanes of their own, next to the stage
                    // tests -- barriers that are NOT yet complete are noticed late by a lone polling lane:
                    // 17.7 instead of 20.1 G attempts/s)
                    const bool whole_part = (C >= 4) && 2 * (c_end - c_begin) < NS;
                    if (!whole_part) {
#pragma unroll 1
                        for (int c = c_begin; c < c_end; ++c) {
                            mbar_wait(&full[stage], fpar);
                            tc::fence_after_sync();
                            const int nt = (dbg & 2) ? 0 : min(CT, Tl - c * CT);
                            const uint64_t adesc0 = tc::make_smem_desc(
                                smem_u32(ring + (size_t)stage * kStageBytes), kALbo, kASbo);
                            const uint32_t d0 = tbase + (uint32_t)(c * CT * NG);
                            if (tc::elect_one()) {
#pragma unroll
                                for (int tt = 0; tt < CT; ++tt) {
                                    if (tt < nt) {
#pragma unroll
                                        for (int p = 0; p < P; ++p)
                                            tc::mma_bf16_ss(d0 + tt * NG,
                                                            adesc0 + (uint64_t)(((tt * P + p) * kTileBytes) >> 4),
                                                            bdesc, IDESC, 1u);
                                    }
                                }
                                tc::mma_commit(&empty[stage]);
                            }
                            __syncwarp();
                            if (++stage == NS) { stage = 0; fpar ^= 1u; }
                        }
                        if (tc::elect_one()) tc::mma_commit(&hdone[kParts * slot + h]);
                        __syncwarp();
                        continue;
                    }
                    if (lane < c_end - c_begin) {
                        int st_l = stage + lane;
                        uint32_t par_l = fpar;
                        if (st_l >= NS) { st_l -= NS; par_l ^= 1u; }
                        mbar_wait(&full[st_l], par_l);
                    }
                    __syncwarp();
                    tc::fence_after_sync();
                    // (what bounds this loop is the tensor pipe's rate per INSTRUCTION: an
                    // M128 x N64 x K16 MMA takes 49 clocks in isolation, 61 with a commit per six
                    // (tools/mma_dep_bench.py; the floor is ~41 clocks however small N is), i.e.
                    // 24 MMAs = ~1.5 k of the clocks this warp needs per block; the copies themselves
                    // cost nothing)
                    if (tc::elect_one()) {
                        int st_c = stage;
#pragma unroll 1
                        for (int c = c_begin; c < c_end; ++c) {
                            const int nt = (dbg & 2) ? 0 : min(CT, Tl - c * CT);
                            const uint64_t adesc0 = tc::make_smem_desc(
                                smem_u32(ring + (size_t)st_c * kStageBytes), kALbo, kASbo);
                            const uint32_t d0 = tbase + (uint32_t)(c * CT * NG);
                            // (tile-major; plane-major -- consecutive MMAs on different accumulator
                            // tiles -- was measured slower, and tools/mma_dep_bench.py shows no
                            // penalty for back-to-back MMAs on one accumulator)
#pragma unroll
                            for (int tt = 0; tt < CT; ++tt) {
                                if (tt < nt) {
#pragma unroll
                                    for (int p = 0; p < P; ++p)
                                        tc::mma_bf16_ss(d0 + tt * NG,
                                                        adesc0 + (uint64_t)(((tt * P + p) * kTileBytes) >> 4),
                                                        bdesc, IDESC, 1u);
                                }
                            }
                            tc::mma_commit(&empty[st_c]);
                            if (++st_c == NS) st_c = 0;
                        }
                        tc::mma_commit(&hdone[kParts * slot + h]);
                    }
                    __syncwarp();
                    stage += c_end - c_begin;
                    if (stage >= NS) { stage -= NS; fpar ^= 1u; }
                }
                SG_STAMP(12);
            }
        }
    }

    // ------------------------------------------------------------ teardown
    tc::fence_before_sync();
    __syncthreads();
    if (a.dbg && tid == 0 && blockIdx.x < 512) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.dbg[8192 + blockIdx.x * 4 + 1] = (long long)gt;
        a.dbg[8192 + blockIdx.x * 4 + 3] = (long long)clock64();
    }
    if (C > 1) cluster_sync_all();   // no CTA exits while a peer may still write into it
    if (warp == 0) tc::tmem_dealloc(tbase, (uint32_t)tmem_cols);
}


// ---------------------------------------------------------------- MMA issue-rate probe
// One thread issues `iters` x 32 back-to-back tcgen05.mma (M = 128, K = 16, bf16) on static
// shared-memory operands, cycling over 32 A tiles and the accumulator columns; reports clocks
// per MMA.  variant: 0 = A MN-major no swizzle (SBO 128), 1 = same with SBO 144,
// 2 = A K-major no swizzle, 3 = A K-major 128B swizzle, 4 = A MN-major 128B swizzle.
__global__ void __launch_bounds__(128, 1)
tc_mma_bench_kernel(int variant_in, int n_dim, int iters, long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t cbar;                  // sink of the periodic commits
    const int commit_every = variant_in >> 4;  // 0 = only the final commit
    const int variant = variant_in & 15;
    __shared__ uint32_t tptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (160 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&cbar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tc::tmem_alloc(&tptr, 512);
        tc::tmem_relinquish();
    }
    fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tptr;
    if (commit_every == 15) {
        // probe: how long does a tcgen05.ld of an untouched TMEM column take while `iters` x 32
        // MMAs (to other columns) are queued?  warp 1 issues the MMAs, sets a flag after the
        // first 32, warp 0 then issues one tcgen05.ld (+ wait) and times it.
        __shared__ volatile int go;
        __shared__ long long t_mma_done;
        if (tid == 0) go = 0;
        __syncthreads();
        if (warp == 1) {
            const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 150 * 1024);
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(16 >> 3) << 17) | (8u << 24);
            const uint64_t bdesc = tc::make_smem_desc(b0, 16 * 16, 128);
            const uint64_t adesc0 = tc::make_smem_desc(a0, 2048, 128);
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    if (tc::elect_one())
                        tc::mma_bf16_ss(tbase + (t & 15) * 16, adesc0 + (uint64_t)((t * 4608) >> 4), bdesc, idesc, 1u);
                }
                if (it == 0 && tid == 32) go = 1;
            }
            if (tc::elect_one()) tc::mma_commit(&bar);
            mbar_wait(&bar, 0);
            if (tid == 32) t_mma_done = clock64() - t0;
        } else if (warp == 0) {
            while (go == 0) {
            }
            float v[16];
            const long long t0 = clock64();
            tc::tmem_ld16(tbase + 400, v);
            tc::wait_ld();
            const long long t1 = clock64();
            if (tid == 0) out[0] = (t1 - t0) + (v[0] == 12345.0f ? 1 : 0);
        }
        __syncthreads();
        if (tid == 0) out[1] = t_mma_done;
        tc::fence_before_sync();
        __syncthreads();
        if (warp == 0) tc::tmem_dealloc(tbase, 512);
        return;
    }
    if (warp == 1) {
        // the whole warp runs the loop (warp-uniform descriptor math); one elected lane issues
        const uint32_t a0 = smem_u32(smem);            // 32 A tiles, 4608 B apart (128 KB + pad)
        const uint32_t b0 = smem_u32(smem + 150 * 1024);
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_dim >> 3) << 17) | (8u << 24);
        uint32_t lbo, sbo, tile = 4608;
        uint64_t lt = 0;
        if (variant == 0) { lbo = 2048; sbo = 128; idesc |= (1u << 15); }
        else if (variant == 1) { lbo = 2304; sbo = 144; idesc |= (1u << 15); }
        else if (variant == 2) { lbo = 128; sbo = 256; }
        else if (variant == 3) { lbo = 16; sbo = 1024; lt = 2; tile = 4096; }
        else { lbo = 1024; sbo = 2048; lt = 2; idesc |= (1u << 15); tile = 4096; }
        // B: K-major no swizzle, n_dim rows: (n/8)*128 + (k/8)*(n_dim*16)
        const uint64_t bdesc = tc::make_smem_desc(b0, (uint32_t)n_dim * 16, 128);
        const int ncol_tiles = 512 / n_dim;
        const uint64_t adesc0 = tc::make_smem_desc(a0, lbo, sbo) | (lt << 61);
        const long long t0 = clock64();
        auto adesc_of = [&](int t) {
            return adesc0 + (uint64_t)((variant == 3 ? ((t >> 2) * 16384 + (t & 3) * 32) : t * tile) >> 4);
        };
        if (commit_every == 0) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    if (tc::elect_one())
                        tc::mma_bf16_ss(tbase + (t & (ncol_tiles - 1)) * n_dim, adesc_of(t), bdesc, idesc, 1u);
                }
            }
        } else if (commit_every == 14) {
            // the fix: one election per MMA, descriptors computed in converged code
            for (int it = 0; it < iters; ++it) {
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    const int t0c = (c * 12) & 31;
                    const uint64_t ad0 = adesc_of(t0c & 16);
                    const uint32_t dd0 = tbase + (uint32_t)((t0c & (ncol_tiles - 1)) * n_dim);
#pragma unroll
                    for (int t = 0; t < 12; ++t) {
                        if (tc::elect_one())
                            tc::mma_bf16_ss(dd0 + (t & 3) * n_dim, ad0 + (uint64_t)((t * 4608) >> 4), bdesc, idesc, 1u);
                    }
                    if (tc::elect_one()) tc::mma_commit(&cbar);
                    __syncwarp();
                }
            }
        } else if (commit_every >= 8 && commit_every <= 11) {
            // 8: commits only (96 per iteration, no MMA); 9: the sweep kernel's ring-stage pattern,
            // plane-major (2 tiles x 3 planes, consecutive MMAs on different accumulators, commit);
            // 10: every MMA accumulates into the SAME tile; 11: the stage pattern tile-major (the
            // three planes of a tile back to back on one accumulator, commit)
            for (int it = 0; it < iters; ++it) {
#pragma unroll 1
                for (int c = 0; c < 16; ++c) {
                    const uint64_t ad0 = adesc_of((c * 6) & 24);
                    if (tc::elect_one()) {
                        if (commit_every == 8) {
#pragma unroll
                            for (int t = 0; t < 6; ++t) tc::mma_commit(&cbar);
                        } else {
#pragma unroll
                            for (int t = 0; t < 6; ++t) {
                                const int tile = (commit_every == 10) ? 0 : (commit_every == 9) ? (t & 1) : (t / 3);
                                tc::mma_bf16_ss(tbase + (uint32_t)(tile * n_dim), ad0 + (uint64_t)((t * 4608) >> 4),
                                                bdesc, idesc, 1u);
                            }
                            tc::mma_commit(&cbar);
                        }
                    }
                    __syncwarp();
                }
            }
        } else if (commit_every == 12 || commit_every == 13) {
            // the sweep kernel's issue pattern: one elected lane issues 12 MMAs (4 tiles x 3 planes)
            // and (12) a commit, then the warp reconverges
            for (int it = 0; it < iters; ++it) {
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    const int t0c = (c * 12) & 31;
                    const uint64_t ad0 = adesc_of(t0c & 16);
                    const uint32_t dd0 = tbase + (uint32_t)((t0c & (ncol_tiles - 1)) * n_dim);
                    if (tc::elect_one()) {
#pragma unroll
                        for (int t = 0; t < 12; ++t)
                            tc::mma_bf16_ss(dd0 + (t & 3) * n_dim, ad0 + (uint64_t)((t * 4608) >> 4), bdesc, idesc, 1u);
                        if (commit_every == 12) tc::mma_commit(&cbar);
                    }
                    __syncwarp();
                }
            }
        } else {
            for (int it = 0; it < iters; ++it) {
#pragma unroll 1
                for (int c = 0; c < 16; ++c) {
                    if (tc::elect_one()) {
#pragma unroll
                        for (int t = 0; t < 6; ++t) {
                            const int tt = (c * 6 + t) & 31;
                            tc::mma_bf16_ss(tbase + (tt & (ncol_tiles - 1)) * n_dim, adesc_of(tt), bdesc, idesc, 1u);
                        }
                        if (commit_every == 6) tc::mma_commit(&cbar);
                    }
                    __syncwarp();
                }
            }
        }
        const long long t1 = clock64();
        if (tc::elect_one()) tc::mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (tid == 32) {
            // variants 12/13 and 6/7 issue 96 MMAs per iteration, the plain loop 32
            const int per_it = commit_every ? 96 : 32;
            out[0] = (t1 - t0) * 32 / per_it;
            out[1] = (t2 - t0) * 32 / per_it;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

}  // namespace

cudaError_t launch_split_planes(const float* Jt, int n, int n_pad, void* Jp, int n_tc, int* used,
                                cudaStream_t st) {
    split_planes_kernel<<<1184, 256, 0, st>>>(Jt, n, n_pad, static_cast<__nv_bfloat16*>(Jp), n_tc, used);
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


cudaError_t launch_tc_mma_bench(int variant, int n_dim, int iters, long long* out, cudaStream_t st) {
    const int smem = 160 * 1024;
    cudaError_t err = cudaFuncSetAttribute(tc_mma_bench_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    tc_mma_bench_kernel<<<1, 128, smem, st>>>(variant, n_dim, iters, out);
    return cudaGetLastError();
}

bool sweep_tc_supported(int n, int n_tc) { return n_tc > 0 && n_tc <= 4096 && n >= 16; }

size_t sweep_tc_sites_bytes(int n, int n_sweeps, int R) {
    const int n_s = (n + 15) / 16 * 16;
    return (((size_t)n_sweeps * n_s * sizeof(uint16_t) + 15) & ~(size_t)15) +
           (size_t)((R + kG - 1) / kG) * sizeof(int);
}

// bytes of operand stream one sweep needs
size_t sweep_tc_stream_bytes_per_sweep(int n, int n_tc, int planes) {
    const int nblk = (n + kBlk - 1) / kBlk;
    const int nchunk = (n_tc / kTileM + kChunkTiles - 1) / kChunkTiles;
    // operand chunks + the decision tables of every block (rounded to 128 B)
    return (size_t)nblk * nchunk * kChunkTiles * planes * kTileBytes +
           (((size_t)nblk * kTabBytes + 127) / 128) * 128;
}

// Sweeps per work item for `groups` replica groups on `n_cta` persistent CTAs: the chunking whose
// most loaded CTA finishes first (an item costs its sweeps plus ~3% of a sweep for moving the
// group's state through HBM); ties go to the longer items.
static int tc_pick_spi(int groups, int S, int n_sm, double* cost_out = nullptr) {
    if (groups <= n_sm || S <= 1) {
        if (cost_out) *cost_out = S + 0.03;
        return S;
    }
    int best_spi = S;
    double best_cost = 1e300;
    for (int spi = S; spi >= 1; --spi) {
        const int chunks = (S + spi - 1) / spi;
        const long long n_items = (long long)groups * chunks;
        const int n_cta = (int)(n_items < n_sm ? n_items : n_sm);
        double worst = 0.0;
        for (int c = 0; c < n_cta; ++c) {
            double load = 0.0;
            for (long long it = c; it < n_items; it += n_cta) {
                const int ch = (int)(it / groups);
                const int len = (ch * spi + spi <= S) ? spi : S - ch * spi;
                load += len + 0.03;
            }
            if (load > worst) worst = load;
        }
        if (worst < best_cost - 1e-9) {
            best_cost = worst;
            best_spi = spi;
        }
    }
    if (cost_out) *cost_out = best_cost;
    return best_spi;
}

template <int P, bool INJ, int C>
static cudaError_t launch_tc_variant(const SweepDev& a, const __nv_bfloat16* J, int n_tc,
                                     const uint16_t* sites, int n_s, int NS, int cols,
                                     const unsigned char* Q, const unsigned char* tabs, int s0, int s1,
                                     int spi, int grid_groups, int* done, int dbg, size_t smem,
                                     cudaStream_t st) {
    cudaError_t err = cudaFuncSetAttribute(sweep_tc_kernel<P, INJ, C>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(grid_groups * C), 1, 1);
    cfg.blockDim = dim3(tc_threads(C), 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (C > 1) ? 1 : 0;
    if (getenv("SG_TC_VERBOSE")) {
        int ncl = -1;
        cudaOccupancyMaxActiveClusters(&ncl, sweep_tc_kernel<P, INJ, C>, &cfg);
        fprintf(stderr, "[sg] sweep_tc C=%d grid=%d smem=%zu NS=%d sweeps/item=%d max active clusters=%d\n",
                C, grid_groups * C, smem, NS, spi, ncl);
    }
    return cudaLaunchKernelEx(&cfg, sweep_tc_kernel<P, INJ, C>, a, J, n_tc, sites, n_s, NS, cols, Q,
                              tabs, s0, s1, spi, done, dbg);
}

// CTAs per replica group the launcher uses for this shape: 4 = a thread-block cluster holds 64
// replicas and each CTA owns a quarter of the field columns (needs n_tc / 128 tiles divisible by 8
// and more than 32 replicas), 2 = a pair holds 32 replicas (same shapes, more than 16 replicas),
// else 1.  SG_TC_CLUSTER=1 / 2 caps it (tests, A/B timing).
int sweep_tc_cluster_size(int n_tc, int R) {
    const int T = n_tc / kTileM;
    int C = 1;
    if (T % (2 * kChunkTiles) == 0) C = (R > 2 * kG) ? 4 : (R > kG) ? 2 : 1;
    if (const char* c_env = getenv("SG_TC_CLUSTER")) {
        const int cap = atoi(c_env);
        if (cap == 1 || cap == 2) C = (C < cap) ? C : cap;
        // clusters of 8 (128 replicas, an eighth of the columns, four decision warps) are built and
        // bit-identical, but slower: 15 clusters fit the GPU (120 SMs) and four decision warps per
        // SM slow each other down (11 G attempts/s at the headline shape against 19 for C = 4)
        if (cap == 8 && C == 4 && R > 4 * kG) C = 8;
    }
    return C;
}

// Launches with the persistent work-item schedule spin-wait on each other's CTAs, so two of them
// must never share the GPU (each could hold the SMs the other's missing CTAs need).  Launches of
// one process are chained with an event per device, whatever streams they are on.
struct TcItemGate {
    std::mutex mu;
    cudaEvent_t ev[64] = {};
    bool have[64] = {};
};
static TcItemGate g_item_gate;

// Clusters of 4 leave SMs idle (33 clusters = 132 of 148 SMs: a GPC holds a whole number of them).
// A second launch of the SAME kernel family with cluster PAIRS runs the last replicas of the engine
// on those SMs at the same time, on a side stream of its own (non-blocking: the caller's stream may
// be the legacy default stream).  Replica r draws the same Philox numbers and sees the same MMAs in
// either form, so the results do not depend on the split.
struct TcSideStream {
    std::mutex mu;
    cudaStream_t st[64] = {};
    cudaEvent_t ready[64] = {}, done[64] = {};
    bool have[64] = {};
};
static TcSideStream g_side;

// (call with g_side.mu held; the caller keeps it until its launches and event calls are enqueued, so
// that two host threads never interleave their records of the shared events)
static cudaError_t tc_side_stream(int dev, cudaStream_t* st, cudaEvent_t* ready, cudaEvent_t* done) {
    if (!g_side.have[dev]) {
        cudaError_t e = cudaStreamCreateWithFlags(&g_side.st[dev], cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
        e = cudaEventCreateWithFlags(&g_side.ready[dev], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        e = cudaEventCreateWithFlags(&g_side.done[dev], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        g_side.have[dev] = true;
    }
    *st = g_side.st[dev];
    *ready = g_side.ready[dev];
    *done = g_side.done[dev];
    return cudaSuccess;
}

// the slice [r0, r0 + count) of the replicas as a launch of its own
static SweepDev tc_slice(const SweepDev& a, int r0, int count) {
    SweepDev b = a;
    b.spins += (size_t)r0 * a.n_pad;
    b.fields += (size_t)r0 * a.n_pad;
    b.energy += r0;
    b.best_energy += r0;
    b.best_spins += (size_t)r0 * a.n_pad;
    b.accepted += r0;
    b.temps += (long long)r0 * a.t_rs;
    if (b.energy_trace) b.energy_trace += r0;
    b.trace_ld = a.trace_ld ? a.trace_ld : a.R;
    b.rep_base = a.rep_base + r0;
    b.R = count;
    if (r0) b.dbg = nullptr;   // (the development timeline follows block 0 of the main launch)
    return b;
}

// clusters of C CTAs that can be resident at the same time (the persistent work-item schedule must
// not launch more: a waiting cluster would otherwise hold the SMs a cluster it depends on needs)
template <int C>
static int tc_max_clusters(size_t smem) {
    cudaFuncSetAttribute(sweep_tc_kernel<3, false, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C, 1, 1);
    cfg.blockDim = dim3(tc_threads(C), 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, sweep_tc_kernel<3, false, C>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return ncl;
}

// The SMs the clusters of 4 leave idle take the last replicas as cluster pairs, in groups of 32: one
// pair of CTAs per group for all sweeps of the launch, WITHOUT the work-item schedule -- the pairs
// never wait for each other, so the two concurrent launches cannot hold SMs the other one's missing
// CTAs need.  The number of groups is chosen so that neither launch waits for the other (a pair
// needs ~1.4 x the time of a cluster of 4 per group and sweep).  Returns the number of replicas for
// the pairs (SG_TC_HYBRID=0: none, SG_TC_HYBRID_M forces the number of groups per idle pair).
static int tc_side_split(int R, int S, int n_slots, int n_sm) {
    const char* hy = getenv("SG_TC_HYBRID");
    const int idle_pairs = (n_sm - 4 * n_slots) / 2;
    const int groups4 = (R + 4 * kG - 1) / (4 * kG);
    if ((hy && atoi(hy) == 0) || idle_pairs < 1 || groups4 <= n_slots) return 0;
    const char* rt = getenv("SG_TC_HYBRID_RATIO");
    const double ratio = rt ? atof(rt) : 1.4;
    // the search below costs ~0.5 ms of host time: remember the answer for the shape (a launch that
    // the host does not run ahead of would otherwise pay it on the device's clock)
    static std::mutex memo_mu;
    static struct { int R, S, n_slots, n_sm, gb; double ratio; } memo[8];
    static int memo_n = 0;
    const bool plain = !getenv("SG_TC_HYBRID_M") && !getenv("SG_TC_HYBRID_GROUPS");
    if (plain) {
        std::lock_guard<std::mutex> lock(memo_mu);
        for (int i = 0; i < memo_n; ++i)
            if (memo[i].R == R && memo[i].S == S && memo[i].n_slots == n_slots && memo[i].n_sm == n_sm &&
                memo[i].ratio == ratio)
                return memo[i].gb * 2 * kG;
    }
    double best = 0.0;
    tc_pick_spi(groups4, S, n_slots, &best);
    int best_gb = 0;
    for (int gb = 1; gb <= 8 * idle_pairs; ++gb) {
        const int rb = gb * 2 * kG;
        if (rb * 4 > R) break;
        const int ga = (R - rb + 4 * kG - 1) / (4 * kG);
        if (ga <= n_slots) break;
        double ta = 0.0;
        tc_pick_spi(ga, S, n_slots, &ta);
        const double tb = ((gb + idle_pairs - 1) / idle_pairs) * (S * ratio + 0.03);
        const double t = ta > tb ? ta : tb;
        if (t < best * 0.985 && (best_gb == 0 || t < best - 1e-9)) {
            best = t;
            best_gb = gb;
        }
    }
    if (plain) {
        std::lock_guard<std::mutex> lock(memo_mu);
        const int i = memo_n < 8 ? memo_n++ : 7;
        memo[i] = {R, S, n_slots, n_sm, best_gb, ratio};
    }
    if (const char* fm = getenv("SG_TC_HYBRID_M")) {
        const int v = atoi(fm);
        if (v >= 0 && v * idle_pairs * 2 * kG * 2 <= R) best_gb = v * idle_pairs;
    }
    if (const char* fg = getenv("SG_TC_HYBRID_GROUPS")) {
        const int v = atoi(fg);
        if (v >= 0 && v * 2 * kG * 2 <= R) best_gb = v;
    }
    return best_gb * 2 * kG;
}

// replicas a launch of n_sweeps sweeps hands to the pairs (0: everything runs on one cluster size)
int sweep_tc_side_replicas(int n, int n_tc, int planes, int R, int n_sweeps) {
    if (planes < 1 || planes > 3 || !sweep_tc_supported(n, n_tc) || sweep_tc_cluster_size(n_tc, R) != 4) return 0;
    if (getenv("SG_TC_SM")) return 0;
    int NS = kMaxStagesTc;
    while (NS > 2 && tc_layout(n_tc, planes, NS, 4).total > 227 * 1024) --NS;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    int n_slots = tc_max_clusters<4>(tc_layout(n_tc, planes, NS, 4).total);
    if (n_slots > n_sm / 4) n_slots = n_sm / 4;
    if (n_slots < 1) return 0;
    return tc_side_split(R, n_sweeps, n_slots, n_sm);
}

cudaError_t launch_sweep_tc(const SweepDev& a, const void* Jp, int n_tc, int planes, bool inject,
                            void* sites_buf, void* stream_buf, size_t stream_cap,
                            uint64_t* launches, KernelTimer* timer, cudaStream_t st) {
    if (planes < 1 || planes > 3 || !sweep_tc_supported(a.n, n_tc)) return cudaErrorInvalidValue;
    if (a.site_mode == 3 || (a.site_mode == 2 && a.s_bs != 0)) return cudaErrorInvalidValue;
    const int n_s = (a.n + 15) / 16 * 16;
    uint16_t* sites = static_cast<uint16_t*>(sites_buf);
    {
        const int total = a.n_sweeps * (n_s / 4);
        int grid = (total + 255) / 256;
        if (grid > 1184) grid = 1184;
        tc_sites_kernel<<<grid, 256, 0, st>>>(a.site_mode, a.seed, a.sweep_base, a.n, n_s,
                                              a.n_sweeps, a.sites, a.s_ss, sites);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        ++*launches;
    }
    // C = 4 / 2 (default where the shape allows it): a cluster per 64 / 32 replicas, each CTA
    // owning 1/C of the field columns -- 1/C of the operand bytes per SM and attempt, bit-identical
    // results (C = 1: one CTA per 16 replicas; SG_TC_CLUSTER=1 / 2 caps C).
    const int T = n_tc / kTileM;
    const int C = sweep_tc_cluster_size(n_tc, a.R);
    int NS = kMaxStagesTc;
    while (NS > 2 && tc_layout(n_tc, planes, NS, C).total > 227 * 1024) --NS;
    if (const char* ns_env = getenv("SG_TC_STAGES")) {
        const int v = atoi(ns_env);
        if (v >= 2 && v <= NS) NS = v;
    }
    const size_t smem = tc_layout(n_tc, planes, NS, C).total;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    int cols = 32;
    while (cols < T * kG) cols *= 2;
    const __nv_bfloat16* J = static_cast<const __nv_bfloat16*>(Jp);
    const int nblk = (a.n + kBlk - 1) / kBlk;
    const int nchunk = (T + kChunkTiles - 1) / kChunkTiles;
    const size_t per_sweep = sweep_tc_stream_bytes_per_sweep(a.n, n_tc, planes);
    const size_t q_per_sweep = (size_t)nblk * nchunk * kChunkTiles * planes * kTileBytes;
    const int sub = (int)(stream_cap / per_sweep);
    if (sub < 1) return cudaErrorInvalidValue;
    // development aid (timing experiments only, results are wrong): SG_TC_DBG bit0 = constant
    // operand chunk, bit1 = no MMA, bit2 = no raw TMEM reads
    const char* dbg_env = getenv("SG_TC_DBG");
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    const char* spi_s = getenv("SG_TC_SPI");
    const int spi_env = spi_s ? atoi(spi_s) : -1;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    bool sm_env_set = false;
    if (const char* sm_env = getenv("SG_TC_SM")) {   // tests: pretend the GPU has fewer SMs
        const int v = atoi(sm_env);
        if (v >= 1 && v < n_sm) {
            n_sm = v;
            sm_env_set = true;
        }
    }
    // per-group progress counters live behind the site tables
    int* done = reinterpret_cast<int*>(static_cast<unsigned char*>(sites_buf) +
                                       (((size_t)a.n_sweeps * n_s * sizeof(uint16_t) + 15) & ~(size_t)15));
    cudaError_t err = cudaSuccess;
    for (int s0 = 0; s0 < a.n_sweeps; s0 += sub) {
        const int s1 = (s0 + sub < a.n_sweeps) ? s0 + sub : a.n_sweeps;
        const size_t units = (size_t)(s1 - s0) * q_per_sweep / 16;
        // persistent CTAs over (sweep chunk, replica group) items when there are more groups than
        // SMs: no partial last wave (SG_TC_SPI=0 forces one CTA per group for all sweeps)
        // CTAs (C = 1) or clusters (C = 2, 4) resident at once
        int n_slots = n_sm;
        if (C > 1) {
            n_slots = (C == 8) ? tc_max_clusters<8>(smem) : (C == 4) ? tc_max_clusters<4>(smem) : tc_max_clusters<2>(smem);
            if (n_slots > n_sm / C) n_slots = n_sm / C;
            if (sm_env_set && n_sm / C < n_slots) n_slots = n_sm / C > 0 ? n_sm / C : 1;
            if (n_slots < 1) n_slots = 1;
        }
        const int r_side = (C == 4 && !inject && !sm_env_set && dev >= 0 && dev < 64)
                               ? tc_side_split(a.R, s1 - s0, n_slots, n_sm) : 0;
        const SweepDev a4 = r_side ? tc_slice(a, 0, a.R - r_side) : a;
        const int groups = (a4.R + kG * C - 1) / (kG * C);
        int spi = s1 - s0, grid_groups = groups;
        bool item_mode = false;
        if (groups > n_slots) {
            spi = tc_pick_spi(groups, s1 - s0, n_slots);
            if (spi_env > 0) spi = spi_env < s1 - s0 ? spi_env : s1 - s0;
            if (spi_env == 0) spi = s1 - s0;
            const int chunks = (s1 - s0 + spi - 1) / spi;
            if (chunks > 1) {
                const long long n_items = (long long)groups * chunks;
                grid_groups = (int)(n_items < n_slots ? n_items : n_slots);
                err = cudaMemsetAsync(done, 0, (size_t)groups * sizeof(int), st);
                if (err != cudaSuccess) return err;
                item_mode = true;
            }
        }
        std::unique_lock<std::mutex> gate(g_item_gate.mu, std::defer_lock);
        if (item_mode && dev >= 0 && dev < 64) {
            gate.lock();
            if (g_item_gate.have[dev]) {
                err = cudaStreamWaitEvent(st, g_item_gate.ev[dev], 0);
                if (err != cudaSuccess) return err;
            }
        }
        cudaStream_t side_st = nullptr;
        cudaEvent_t side_ready = nullptr, side_done = nullptr;
        SweepDev a2 = a;
        int NS2 = kMaxStagesTc, grid2 = 0, spi2 = 0;
        size_t smem2 = 0;
        std::unique_lock<std::mutex> side_lock(g_side.mu, std::defer_lock);
        if (r_side) {
            side_lock.lock();
            err = tc_side_stream(dev, &side_st, &side_ready, &side_done);
            if (err != cudaSuccess) return err;
            a2 = tc_slice(a, a.R - r_side, r_side);
            while (NS2 > 2 && tc_layout(n_tc, planes, NS2, 2).total > 227 * 1024) --NS2;
            smem2 = tc_layout(n_tc, planes, NS2, 2).total;
            if (smem2 > 227 * 1024) return cudaErrorInvalidValue;
            grid2 = r_side / (2 * kG);
            spi2 = s1 - s0;
        }
        unsigned char* tabs = static_cast<unsigned char*>(stream_buf) + (size_t)sub * q_per_sweep;
        const unsigned char* Qc = static_cast<const unsigned char*>(stream_buf);
        int ggrid = (int)((units + 255) / 256 < (size_t)148 * 16 ? (units + 255) / 256 : (size_t)148 * 16);
#define SG_TC(P, INJ)                                                                          \
    {                                                                                          \
        if (timer) timer->begin(1, st);                                                        \
        tc_gather_kernel<P><<<ggrid, 256, 0, st>>>(J, a.n, n_tc, sites, n_s, s0, s1 - s0, nblk,\
                                                   nchunk, static_cast<uint4*>(stream_buf));   \
        tc_tables_kernel<P><<<(s1 - s0) * nblk, 256, 0, st>>>(J, a.n, n_tc, sites, n_s, s0,    \
                                                              nblk, tabs);                     \
        if (timer) timer->end(st);                                                             \
        err = cudaGetLastError();                                                              \
        if (err != cudaSuccess) return err;                                                    \
        if (timer) timer->begin(0, st);                                                        \
        if (r_side) {                                                                          \
            err = cudaEventRecord(side_ready, st);                                             \
            if (err != cudaSuccess) return err;                                                \
        }                                                                                      \
        err = (C == 8) ? launch_tc_variant<P, INJ, 8>(a4, J, n_tc, sites, n_s, NS, cols, Qc, tabs,\
                                                      s0, s1, spi, grid_groups, done, dbg, smem, st) \
            : (C == 4) ? launch_tc_variant<P, INJ, 4>(a4, J, n_tc, sites, n_s, NS, cols, Qc, tabs,\
                                                      s0, s1, spi, grid_groups, done, dbg, smem, st) \
            : (C == 2) ? launch_tc_variant<P, INJ, 2>(a4, J, n_tc, sites, n_s, NS, cols, Qc, tabs,\
                                                      s0, s1, spi, grid_groups, done, dbg, smem, st) \
                       : launch_tc_variant<P, INJ, 1>(a4, J, n_tc, sites, n_s, NS, cols, Qc, tabs,\
                                                      s0, s1, spi, grid_groups, done, dbg, smem, st); \
        if (err != cudaSuccess) return err;                                                    \
        if (r_side) {   /* the pairs: one group of 32 replicas each for all sweeps, no item waits */ \
            err = cudaStreamWaitEvent(side_st, side_ready, 0);                                 \
            if (err != cudaSuccess) return err;                                                \
            err = launch_tc_variant<P, INJ, 2>(a2, J, n_tc, sites, n_s, NS2, cols, Qc, tabs, s0, s1,\
                                               spi2, grid2, done + groups, dbg, smem2, side_st);\
            if (err != cudaSuccess) return err;                                                \
            err = cudaEventRecord(side_done, side_st);                                         \
            if (err != cudaSuccess) return err;                                                \
            err = cudaStreamWaitEvent(st, side_done, 0);                                       \
            if (err != cudaSuccess) return err;                                                \
            ++*launches;                                                                       \
        }                                                                                      \
        if (timer) timer->end(st);                                                             \
    }
        if (inject) {
            if (planes == 1) SG_TC(1, true) else if (planes == 2) SG_TC(2, true) else SG_TC(3, true)
        } else {
            if (planes == 1) SG_TC(1, false) else if (planes == 2) SG_TC(2, false) else SG_TC(3, false)
        }
#undef SG_TC
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        if (gate.owns_lock()) {
            if (!g_item_gate.have[dev]) {
                err = cudaEventCreateWithFlags(&g_item_gate.ev[dev], cudaEventDisableTiming);
                if (err != cudaSuccess) return err;
                g_item_gate.have[dev] = true;
            }
            err = cudaEventRecord(g_item_gate.ev[dev], st);
            if (err != cudaSuccess) return err;
        }
        *launches += 3;
    }
    return err;
}

}  // namespace sg
