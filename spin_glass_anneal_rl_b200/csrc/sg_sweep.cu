// sg_sweep.cu -- K1: the replica-batched Monte Carlo sweep for sm_100a.
//
// Replaces SpinDynamics.sweep() / _metropolis_update / _glauber_update /
// _heat_bath_update (reference core/spin_dynamics.py:61-94, 131-191) and the
// never-launched metropolis_update_kernel (annealing/cuda_kernels.py:8-50).
//
// One thread block per SM, 8 warps (2 per SM sub-partition, up to 255 registers):
//
//   * 7 BULK warps (224 threads) keep the local fields f[r][j] = h_j + sum_i J_ji s_ri
//     of the block's G replicas RESIDENT IN REGISTERS: thread t owns the columns
//     {896k + 4t + e}.  A block visits the sites of a sweep in ONE order shared by its
//     replicas, so row Jt[site][:] is fetched once per attempt and serves all G
//     replicas: rows are streamed L2/HBM -> shared memory by TMA bulk copies
//     (cp.async.bulk + mbarrier) through a D-stage ring; the last bulk warp to
//     finish an attempt (shared-memory counter) re-arms that stage with the row of
//     attempt +D.  An accepted flip of spin i in replica r is the rank-1 update
//     f[r][:] += -2 s_ri * Jt[i][:]  (packed FFMA2, two columns per instruction;
//     replicas are processed in groups of 4 with delta 0 for the ones that did not
//     flip, so there is one uniform branch per group instead of one per replica).
//
//   * 1 DECISION warp (lane = replica) makes the accept decisions for a whole block
//     of kB = 16 consecutive attempts at a time, entirely in registers, with no
//     inter-warp hand-off per attempt.  The bulk threads publish, two blocks ahead,
//     the raw field values of the 16 sites of a block (as of the end of block k-2)
//     plus two 16x16 tables of couplings among the sites of blocks k-1 and k.  The
//     decision warp brings the 16 values up to date itself: first the flips of block
//     k-1 (cross table), then, attempt by attempt, the flips it has just decided
//     (in-block table) -- the same FMAs, in the same order, that the owning bulk
//     thread executes later, so the value it compares is bit-identical to the
//     sequential algorithm.  Decisions are published per block as ready-made float
//     deltas (-2 s, or 0) and flip masks.  The sequential chain per attempt is ~10
//     dependent instructions of ONE warp instead of a round trip through barriers.
//     (v1: __syncthreads per attempt, ~1500 clk/attempt.  v2-v4: decision ring
//     between owner warps, ~900 clk/attempt floor.  A 9th warp would cap every thread
//     at 168 registers because 3 warps then share one sub-partition.)
//
//   * Thresholds: accept <=> dE < -T ln(u).  The u's come from Philox4x32-10 keyed
//     on (replica, absolute sweep, attempt) and are produced a batch of 32 attempts
//     ahead by the bulk warps.  In injected mode the caller supplies the uniforms and
//     the reference's own comparison u < exp(float(-dE/T)) is evaluated.
//   * Spins live as bit planes in shared memory (decision lane r owns plane r).
//   * After every sweep (block barrier) the energy of each replica is reduced from
//     the resident fields, E = -1/2 sum_j s_j (f_j + h_j), and the best
//     configuration is kept.
#include "sg_common.cuh"
#include "sg_internal.h"

// development aid: -DSG_TIMELINE=1 compiles clock64() stamps into block 0 (tools/timeline.py)
#ifndef SG_TIMELINE
#define SG_TIMELINE 0
#endif
#define SG_TL (SG_TIMELINE != 0)

namespace sg {

namespace {

constexpr int kBulkWarps = kBulkThreads / 32;  // 7
constexpr int kDecWarp = kBulkWarps;           // warp 7
constexpr int kSB = 32;                        // attempts per threshold batch
constexpr int kThetaBufs = 3;
constexpr int kB = 16;                         // attempts per decision block
constexpr int kSlots = 4;                      // decision-block ring
constexpr int kMaxStages = 8;

template <int C, int CPT, int G>
__device__ __forceinline__ void publish_one(const float2 (&f)[G][CPT / 2], float* dst) {
    if constexpr (C < CPT) {
#pragma unroll
        for (int r = 0; r < G; ++r) dst[r] = (C & 1) ? f[r][C / 2].y : f[r][C / 2].x;
    }
}

// the owner thread of a site copies its G field values of local column c to shared memory
// (a switch, so that every register index is a compile-time constant)
template <int CPT, int G>
__device__ __forceinline__ void publish_column(const float2 (&f)[G][CPT / 2], int c, float* dst) {
#define SG_CASE(C) case C: publish_one<C, CPT, G>(f, dst); break;
    switch (c) {
        SG_CASE(0) SG_CASE(1) SG_CASE(2) SG_CASE(3) SG_CASE(4) SG_CASE(5) SG_CASE(6) SG_CASE(7)
        SG_CASE(8) SG_CASE(9) SG_CASE(10) SG_CASE(11) SG_CASE(12) SG_CASE(13) SG_CASE(14)
        SG_CASE(15) SG_CASE(16) SG_CASE(17) SG_CASE(18) SG_CASE(19) SG_CASE(20) SG_CASE(21)
        SG_CASE(22) SG_CASE(23) SG_CASE(24) SG_CASE(25) SG_CASE(26) SG_CASE(27) SG_CASE(28)
        SG_CASE(29) SG_CASE(30) SG_CASE(31)
        default: break;
    }
#undef SG_CASE
}

struct SmemLayout {
    size_t jring, sites, owner, sbits, theta, raw, dec, cin, ccr, amk, red, flags, cnt, mbar, total;
};

__host__ __device__ inline SmemLayout make_layout(int n_pad, int G, int D) {
    SmemLayout L;
    size_t off = 0;
    L.jring = off; off += (size_t)D * n_pad * sizeof(float);
    L.sites = off; off += 2 * (size_t)n_pad * sizeof(uint16_t);
    L.owner = off; off += 2 * (size_t)n_pad * sizeof(uint16_t);
    L.sbits = off; off += (size_t)G * (n_pad / 32) * sizeof(uint32_t);
    L.theta = off; off += (size_t)kThetaBufs * kSB * 32 * sizeof(float);
    L.raw = off;   off += (size_t)kSlots * kB * 32 * sizeof(float);
    L.dec = off;   off += (size_t)kSlots * kB * 32 * sizeof(float);
    L.cin = off;   off += (size_t)kSlots * kB * kB * sizeof(float);
    L.ccr = off;   off += (size_t)kSlots * kB * kB * sizeof(float);
    L.amk = off;   off += (size_t)kSlots * kB * sizeof(uint32_t);
    L.red = off;   off += 8 * 32 * sizeof(float);
    L.flags = off; off += 4 * sizeof(uint32_t);
    L.cnt = off;   off += kMaxStages * sizeof(uint32_t);
    off = (off + 15) & ~(size_t)15;
    L.mbar = off;  off += (kMaxStages + 2 * kSlots) * sizeof(uint64_t);
    L.total = off;
    return L;
}

template <int CPT, int G, bool INJECT>
__global__ void __launch_bounds__(kSweepThreads, 1) sweep_kernel(const SweepDev a) {
    constexpr int KCH = CPT / 4;
    constexpr int NQ = (G + 3) / 4;  // replica groups of 4
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = a.n, n_pad = a.n_pad, D = a.D;
    const int W = n_pad >> 5;  // spin-bit words per replica
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_dec = (warp == kDecWarp);

    const SmemLayout L = make_layout(n_pad, G, D);
    float* Jring = reinterpret_cast<float*>(smem + L.jring);
    uint16_t* sites_s = reinterpret_cast<uint16_t*>(smem + L.sites);
    uint16_t* owner_s = reinterpret_cast<uint16_t*>(smem + L.owner);  // (thread << 5) | column
    uint32_t* sbits = reinterpret_cast<uint32_t*>(smem + L.sbits);
    float* theta = reinterpret_cast<float*>(smem + L.theta);
    float* raw_s = reinterpret_cast<float*>(smem + L.raw);   // [slot][b][lane]
    float* dec_s = reinterpret_cast<float*>(smem + L.dec);   // [slot][b][lane] field deltas
    float* cin_s = reinterpret_cast<float*>(smem + L.cin);   // [slot][a][b] in-block couplings
    float* ccr_s = reinterpret_cast<float*>(smem + L.ccr);   // [slot][a][b] previous block -> this
    uint32_t* amk_s = reinterpret_cast<uint32_t*>(smem + L.amk);  // [slot][b] flip masks
    float* red = reinterpret_cast<float*>(smem + L.red);
    uint32_t* flags = reinterpret_cast<uint32_t*>(smem + L.flags);
    uint32_t* cnt_s = reinterpret_cast<uint32_t*>(smem + L.cnt);  // warps done with a ring stage
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.mbar);
    uint64_t* rawbar = full + kMaxStages;   // block inputs published (7 bulk warps arrive)
    uint64_t* decbar = rawbar + kSlots;     // block decided (decision warp arrives)

    const int rep0 = blockIdx.x * a.G;
    const int g_act = min(a.G, a.R - rep0);
    const int n_sweeps = a.n_sweeps;
    const int total = n_sweeps * n;  // attempts per replica in this launch (host checks < 2^31)
    const int nblk = (n + kB - 1) / kB;   // decision blocks per sweep
    const int nbat = (n + kSB - 1) / kSB; // threshold batches per sweep
    const uint32_t row_bytes = (uint32_t)n_pad * 4u;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));

    // ------------------------------------------------------------ helpers
    // site table (+ owner thread / local column of every site) of launch-local sweep s
    auto gen_sites = [&](int s) {
        uint16_t* tab = sites_s + (size_t)(s & 1) * n_pad;
        uint16_t* own = owner_s + (size_t)(s & 1) * n_pad;
        auto put = [&](int i, int site) {
            tab[i] = (uint16_t)site;
            const int chunk = site / kColQuantum, rem = site - chunk * kColQuantum;
            own[i] = (uint16_t)(((rem >> 2) << 5) | (chunk * 4 + (rem & 3)));
        };
        if (a.site_mode == 0) {
            for (int i = tid; i < n; i += kSweepThreads) put(i, i);
        } else if (a.site_mode == 1 || a.site_mode == 3) {
            // mode 1: one order for the whole grid; mode 3: an independent order per block
            const unsigned long long sa = a.sweep_base + (unsigned long long)s;
            const uint32_t salt = (a.site_mode == 3) ? (uint32_t)blockIdx.x << 12 : 0u;
            for (int q = tid; q * 4 < n; q += kSweepThreads) {
                const uint4 x = philox4x32_10(
                    make_uint4(kSiteStreamTag, (uint32_t)sa, (uint32_t)(sa >> 32),
                               (uint32_t)q + salt), key);
                const uint32_t v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (q * 4 + e < n) put(q * 4 + e, (int)(v[e] % (uint32_t)n));
            }
        } else {
            const int* src = a.sites + (long long)blockIdx.x * a.s_bs + (long long)s * a.s_ss;
            for (int i = tid; i < n; i += kSweepThreads) put(i, src[i]);
        }
    };

    // thresholds of batch bi (32 attempts) of sweep s into buffer `buf`:
    // quad q covers attempts 4q..4q+3 of the batch, lane = replica
    auto gen_theta = [&](int s, int bi, int buf, int q) {
        const int r = lane;
        const int i0 = bi * kSB + q * 4;
        if (r >= g_act || i0 >= n) return;
        const int rep = rep0 + r;
        const unsigned long long sa = a.sweep_base + (unsigned long long)s;
        const float T = (float)a.temps[(long long)s * a.t_ss + (long long)rep * a.t_rs];
        const uint4 x = philox4x32_10(
            make_uint4((uint32_t)(rep + a.rep_base), (uint32_t)sa, (uint32_t)(sa >> 32), (uint32_t)(i0 >> 2)), key);
        const uint32_t v[4] = {x.x, x.y, x.z, x.w};
        float* dst = theta + (size_t)buf * (kSB * 32) + (q * 4) * 32 + r;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float u = u01(v[e]);
            float th;
            if (a.rule == 0) {
                th = -__logf(u) * T;  // accept <=> dE < -T ln u
            } else {
                th = 0.5f * T * (__logf(u) - __logf(1.0f - u));  // spin up <=> field > th
            }
            dst[e * 32] = th;
        }
    };

    // ------------------------------------------------------------ prologue
    if (tid == 0) {
        for (int d = 0; d < D; ++d) {
            mbar_init(&full[d], 1);
            cnt_s[d] = 0u;
        }
        for (int d = 0; d < kSlots; ++d) {
            mbar_init(&rawbar[d], kBulkWarps);
            mbar_init(&decbar[d], 1);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    gen_sites(0);
    if (n_sweeps > 1) gen_sites(1);
    // spin bit planes from int8 spins
    for (int w = tid; w < G * W; w += kSweepThreads) {
        const int r = w / W, word = w - r * W;
        uint32_t bits = 0;
        if (r < g_act) {
            const uint4* src =
                reinterpret_cast<const uint4*>(a.spins + (size_t)(rep0 + r) * n_pad + word * 32);
            const uint4 lo = src[0], hi = src[1];
            const uint32_t x[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t up = (~x[j]) & 0x80808080u;  // byte >= 0  <=> spin up
                const uint32_t nib =
                    ((up >> 7) & 1u) | ((up >> 14) & 2u) | ((up >> 21) & 4u) | ((up >> 28) & 8u);
                bits |= nib << (4 * j);
            }
        }
        sbits[w] = bits;
    }

    // resident local fields (bulk threads), packed pairs of adjacent columns (FFMA2 operands)
    float2 f[G][CPT / 2];
    if (!is_dec) {
#pragma unroll
        for (int r = 0; r < G; ++r) {
            if (r < g_act) {
                const float4* src =
                    reinterpret_cast<const float4*>(a.fields + (size_t)(rep0 + r) * n_pad);
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    const float4 v = src[k * kBulkThreads + tid];
                    f[r][2 * k + 0] = make_float2(v.x, v.y);
                    f[r][2 * k + 1] = make_float2(v.z, v.w);
                }
            } else {
#pragma unroll
                for (int c = 0; c < CPT / 2; ++c) f[r][c] = make_float2(0.0f, 0.0f);
            }
        }
    }

    // per-replica scalars live in the lanes of the decision warp
    float best_e = 3.0e38f, cur_e = 0.0f;
    unsigned int n_acc = 0;
    if (is_dec && lane < g_act) {
        cur_e = a.energy[rep0 + lane];
        best_e = a.track_best ? a.best_energy[rep0 + lane] : 3.0e38f;
    }
    __syncthreads();

    // row of launch-local attempt (ps, pi) -> stage
    auto issue_row = [&](int stage, int ps, int pi) {
        const int site = sites_s[(size_t)(ps & 1) * n_pad + pi];
        mbar_arrive_expect_tx(&full[stage], row_bytes);
        bulk_g2s(Jring + (size_t)stage * n_pad, a.Jt + (size_t)site * n_pad, row_bytes,
                 &full[stage]);
    };
    if (tid == 0) {
        int ps = 0, pi = 0;
        for (int d = 0; d < D && d < total; ++d) {
            issue_row(d, ps, pi);
            if (++pi == n) { pi = 0; ++ps; }
        }
    }

    // ------------------------------------------------------------ sweeps
    int g = 0;             // launch-local attempt counter (bulk warps)
    int stage = 0;         // TMA ring position
    uint32_t parity = 0;
    int kg = 0;            // launch-global decision-block counter (slot = kg & 3)
    int tbuf0 = 0;         // theta buffer of batch 0 of the current sweep
    long long tw = 0, tu = 0, tr = 0;  // debug: wait / update / release clocks

#pragma unroll 1
    for (int s = 0; s < n_sweeps; ++s) {
        const uint16_t* tab = sites_s + (size_t)(s & 1) * n_pad;
        const uint16_t* own = owner_s + (size_t)(s & 1) * n_pad;
        if (s >= 1 && s + 1 < n_sweeps) gen_sites(s + 1);
        if (!INJECT) gen_theta(s, 0, tbuf0, warp);  // batch 0: 8 warps x 4 attempts
        __syncthreads();

        if (is_dec) {
            // ======================================================== DECISION WARP
            double dT = 1.0;
            if (INJECT && lane < g_act)
                dT = a.temps[(long long)s * a.t_ss + (long long)(rep0 + lane) * a.t_rs];
            int tbuf = tbuf0;  // theta buffer of the batch the current block belongs to
#pragma unroll 1
            for (int k = 0; k < nblk; ++k, ++kg) {
                const int slot = kg & (kSlots - 1);
                const uint32_t par = (uint32_t)(kg >> 2) & 1u;
                const int i0 = k * kB;
                const int nbk = min(kB, n - i0);
                const float* rawp = raw_s + (size_t)slot * kB * 32 + lane;
                float* decp = dec_s + (size_t)slot * kB * 32 + lane;
                const float* cin = cin_s + (size_t)slot * kB * kB;
                const float* ccr = ccr_s + (size_t)slot * kB * kB;
                uint32_t* amk = amk_s + slot * kB;
                if (k > 0 && (k & 1) == 0) tbuf = (tbuf + 1 == kThetaBufs) ? 0 : tbuf + 1;
                const float* thp = theta + (size_t)tbuf * (kSB * 32) + (i0 & 31) * 32 + lane;

                float uu[kB];
                if (INJECT) {
#pragma unroll
                    for (int b = 0; b < kB; ++b)
                        uu[b] = (lane < g_act && b < nbk)
                                    ? a.uniforms[((size_t)(rep0 + lane) * n_sweeps + s) * n + i0 + b]
                                    : 0.0f;
                }
                if (SG_TL && a.dbg && blockIdx.x == 0 && lane == 0 && kg < 256) a.dbg[kg * 8 + 1] = clock64();
                mbar_wait(&rawbar[slot], par);  // raw values + coupling tables of this block
                if (SG_TL && a.dbg && blockIdx.x == 0 && lane == 0 && kg < 256) a.dbg[kg * 8 + 0] = clock64();
                float v[kB];
#pragma unroll
                for (int b = 0; b < kB; ++b) v[b] = rawp[b * 32];

                // flips of the previous block (decided by this warp, not yet in the raw values)
                if (k > 0) {
                    const int ps = (kg - 1) & (kSlots - 1);
                    const float* pdec = dec_s + (size_t)ps * kB * 32 + lane;
                    const uint32_t* pam = amk_s + ps * kB;
#pragma unroll
                    for (int aa = 0; aa < kB; ++aa) {
                        if (pam[aa] != 0u) {
                            const float da = pdec[aa * 32];
                            const float4* row = reinterpret_cast<const float4*>(ccr + aa * kB);
#pragma unroll
                            for (int b4 = 0; b4 < kB / 4; ++b4) {
                                const float4 c4 = row[b4];
                                v[4 * b4 + 0] = fmaf(da, c4.x, v[4 * b4 + 0]);
                                v[4 * b4 + 1] = fmaf(da, c4.y, v[4 * b4 + 1]);
                                v[4 * b4 + 2] = fmaf(da, c4.z, v[4 * b4 + 2]);
                                v[4 * b4 + 3] = fmaf(da, c4.w, v[4 * b4 + 3]);
                            }
                        }
                    }
                }

                // the 16 attempts of this block, strictly in order
#pragma unroll
                for (int aa = 0; aa < kB; ++aa) {
                    bool flip = false, up = false;
                    if (aa < nbk && lane < g_act) {
                        const int site = tab[i0 + aa];
                        uint32_t* wp = &sbits[lane * W + (site >> 5)];
                        const uint32_t w = *wp;
                        up = (w >> (site & 31)) & 1u;
                        const float fv = v[aa];
                        if (!INJECT) {
                            const float th = thp[aa * 32];
                            if (a.rule == 0) {
                                const float x = up ? 2.0f * fv : -2.0f * fv;  // dE = 2 s f
                                flip = x < th;
                            } else {
                                flip = ((fv > th) != up);
                            }
                        } else {
                            if (a.rule == 0) {
                                const float x = up ? 2.0f * fv : -2.0f * fv;
                                // reference: dE <= 0 accepts without a draw; else u < exp(float(-dE/T))
                                flip = (x <= 0.0f) || (uu[aa] < expf((float)(-(double)x / dT)));
                            } else {
                                const float arg = (a.rule == 1) ? (float)(-2.0 * (double)fv / dT)
                                                                : (float)(-2.0 * (1.0 / dT) * (double)fv);
                                const float p_up = 1.0f / (1.0f + expf(arg));
                                flip = ((uu[aa] < p_up) != up);
                            }
                        }
                        if (flip) {
                            *wp = w ^ (1u << (site & 31));
                            ++n_acc;
                            if (a.site_de)   // operator output: accepted dE accumulated per site
                                a.site_de[(size_t)(rep0 + lane) * n + site] += up ? 2.0f * fv : -2.0f * fv;
                        }
                    }
                    const float da = flip ? (up ? -2.0f : 2.0f) : 0.0f;
                    decp[aa * 32] = da;
                    const uint32_t am = __ballot_sync(0xFFFFFFFFu, flip);
                    if (lane == 0) amk[aa] = am;
                    if (am != 0u && aa + 1 < kB) {
                        // bring the later sites of the block up to date (row aa of the in-block table)
                        const float* row = cin + aa * kB;
#pragma unroll
                        for (int b = aa + 1; b < kB; ++b) v[b] = fmaf(da, row[b], v[b]);
                    }
                }
                __syncwarp();  // every lane's stores precede the release below
                if (SG_TL && a.dbg && blockIdx.x == 0 && lane == 0 && kg < 256) a.dbg[kg * 8 + 2] = clock64();
                if (lane == 0) mbar_arrive(&decbar[slot]);
            }
        } else {
            // ======================================================== BULK WARPS
            // publish the inputs of decision block kb (launch-global index kgb) of this sweep
            auto publish_block = [&](int kb, int kgb) {
                const int slot = kgb & (kSlots - 1);
                const int i0 = kb * kB;
                const int nbk = min(kB, n - i0);
                float* rawb = raw_s + (size_t)slot * kB * 32;
                for (int b = 0; b < nbk; ++b) {
                    const int o = own[i0 + b];
                    if ((o >> 5) == tid) publish_column<CPT, G>(f, o & 31, rawb + b * 32);
                }
                // coupling tables: cin[a][b] = Jt[site_a][site_b] (this block), ccr[a][b] =
                // Jt[site_a of the previous block][site_b of this block]
                for (int idx = tid; idx < 2 * kB * kB; idx += kBulkThreads) {
                    const int which = idx >> 8, aa = (idx >> 4) & (kB - 1), b = idx & (kB - 1);
                    float val = 0.0f;
                    if (b < nbk) {
                        const int sb = tab[i0 + b];
                        if (which == 0) {
                            if (aa < nbk) val = a.Jt[(size_t)tab[i0 + aa] * n_pad + sb];
                        } else if (kb > 0) {
                            val = a.Jt[(size_t)tab[i0 - kB + aa] * n_pad + sb];
                        }
                    }
                    (which == 0 ? cin_s : ccr_s)[(size_t)slot * kB * kB + aa * kB + b] = val;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&rawbar[slot]);
            };

            publish_block(0, kg);
            if (nblk > 1) publish_block(1, kg + 1);

#pragma unroll 1
            for (int k = 0; k < nblk; ++k, ++kg) {
                const int slot = kg & (kSlots - 1);
                const int i0 = k * kB;
                const int nbk = min(kB, n - i0);
                // thresholds two blocks (one batch) ahead of the decision warp
                if (!INJECT && (k & 1) == 0) {
                    const int bi = (k >> 1) + 1;
                    if (bi < nbat) {
                        const int buf = (tbuf0 + bi) % kThetaBufs;
                        gen_theta(s, bi, buf, warp);
                        if (warp == 0) gen_theta(s, bi, buf, 7);
                    }
                }
                if (SG_TL && a.dbg && blockIdx.x == 0 && tid == 0 && kg < 256) a.dbg[kg * 8 + 3] = clock64();
                mbar_wait(&decbar[slot], (uint32_t)(kg >> 2) & 1u);  // block k decided
                if (SG_TL && a.dbg && blockIdx.x == 0 && tid == 0 && kg < 256) a.dbg[kg * 8 + 4] = clock64();
                const uint32_t* amk = amk_s + slot * kB;
                const float* decb = dec_s + (size_t)slot * kB * 32;
#pragma unroll 1
                for (int aa = 0; aa < nbk; ++aa) {
                    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                    if (SG_TL && a.dbg) c0 = clock64();
                    const uint32_t am = amk[aa];
                    // Always wait for the row, even when no replica flipped: this is what keeps a
                    // fast warp from lapping the ring (a stage is re-armed only after all 7 bulk
                    // warps released it, and its barrier completes only after that).
                    mbar_wait(&full[stage], parity);
                    if (SG_TL && a.dbg) c1 = clock64();
                    if (am != 0u) {
                        const float4* Jr4 = reinterpret_cast<const float4*>(Jring + (size_t)stage * n_pad);
                        const float4* d4 = reinterpret_cast<const float4*>(decb + aa * 32);
                        float4 jv[KCH];
#pragma unroll
                        for (int kk = 0; kk < KCH; ++kk) jv[kk] = Jr4[kk * kBulkThreads + tid];
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            if ((am >> (4 * q)) & 0xFu) {
                                const float4 dq = d4[q];
                                const float dd[4] = {dq.x, dq.y, dq.z, dq.w};
#pragma unroll
                                for (int rr = 0; rr < 4; ++rr) {
                                    const int r = 4 * q + rr;
                                    if (r < G) {
                                        const float2 d2 = make_float2(dd[rr], dd[rr]);
#pragma unroll
                                        for (int kk = 0; kk < KCH; ++kk) {
                                            f[r][2 * kk + 0] = __ffma2_rn(
                                                d2, make_float2(jv[kk].x, jv[kk].y), f[r][2 * kk + 0]);
                                            f[r][2 * kk + 1] = __ffma2_rn(
                                                d2, make_float2(jv[kk].z, jv[kk].w), f[r][2 * kk + 1]);
                                        }
                                    }
                                }
                            }
                        }
                    }
                    if (SG_TL && a.dbg) c2 = clock64();
                    // release the stage: the LAST bulk warp to finish this attempt re-arms the
                    // stage at once with the row of attempt g + D (rows are D attempts ahead)
                    if (lane == 0) {
                        const uint32_t done = atom_add_acq_rel_shared(&cnt_s[stage], 1u);
                        if (done == (uint32_t)(kBulkWarps - 1)) {
                            cnt_s[stage] = 0u;
                            if (g + D < total) {
                                int pi = i0 + aa + D, ps = s;
                                if (pi >= n) { pi -= n; ++ps; }
                                issue_row(stage, ps, pi);
                            }
                        }
                    }
                    if (SG_TL && a.dbg) { c3 = clock64(); tw += c1 - c0; tu += c2 - c1; tr += c3 - c2; }
                    ++g;
                    if (++stage == D) { stage = 0; parity ^= 1u; }
                }
                if (SG_TL && a.dbg && blockIdx.x == 0 && tid == 0 && kg < 256) a.dbg[kg * 8 + 5] = clock64();
                if (k + 2 < nblk) publish_block(k + 2, kg + 2);
                if (SG_TL && a.dbg && blockIdx.x == 0 && tid == 0 && kg < 256) a.dbg[kg * 8 + 6] = clock64();
            }
        }
        tbuf0 = (tbuf0 + nbat) % kThetaBufs;

        // ---- end of sweep: energies from the resident fields, best tracking
        __syncthreads();
        if (!is_dec) {
            const float4* h4 = reinterpret_cast<const float4*>(a.h);
            float hv[CPT];
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                const float4 v = h4[k * kBulkThreads + tid];
                hv[4 * k + 0] = v.x; hv[4 * k + 1] = v.y; hv[4 * k + 2] = v.z; hv[4 * k + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < G; ++r) {
                float part = 0.0f;
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    const uint32_t nib =
                        (sbits[r * W + k * (kBulkThreads / 8) + (tid >> 3)] >> ((tid & 7) * 4)) & 0xFu;
                    const float fe[4] = {f[r][2 * k].x, f[r][2 * k].y, f[r][2 * k + 1].x,
                                         f[r][2 * k + 1].y};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float t = fe[e] + hv[4 * k + e];
                        part += ((nib >> e) & 1u) ? t : -t;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
                if (lane == 0) red[warp * 32 + r] = part;
            }
        }
        __syncthreads();
        if (is_dec) {
            bool improved = false;
            if (lane < g_act) {
                float acc = 0.0f;
#pragma unroll
                for (int w = 0; w < kBulkWarps; ++w) acc += red[w * 32 + lane];
                cur_e = -0.5f * acc;
                if (a.energy_trace) a.energy_trace[(size_t)s * a.R + rep0 + lane] = cur_e;
                if (a.track_best && cur_e < best_e) {
                    best_e = cur_e;
                    improved = true;
                }
            }
            const uint32_t im = __ballot_sync(0xFFFFFFFFu, improved);
            if (lane == 0) flags[0] = im;
        }
        __syncthreads();
        const uint32_t im = flags[0];
        if (im != 0u && !is_dec) {
#pragma unroll
            for (int r = 0; r < G; ++r) {
                if (im & (1u << r)) {
                    uint32_t* dst =
                        reinterpret_cast<uint32_t*>(a.best_spins + (size_t)(rep0 + r) * n_pad);
#pragma unroll
                    for (int k = 0; k < KCH; ++k) {
                        const uint32_t nib =
                            (sbits[r * W + k * (kBulkThreads / 8) + (tid >> 3)] >> ((tid & 7) * 4)) & 0xFu;
                        uint32_t bytes = 0;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            bytes |= (((nib >> e) & 1u) ? 0x01u : 0xFFu) << (8 * e);
                        dst[k * kBulkThreads + tid] = bytes;
                    }
                }
            }
        }
        __syncthreads();  // bit planes / flags are modified again by the next sweep
    }

    if (SG_TL && a.dbg && blockIdx.x == 0 && (tid == 0 || tid == 96)) {
        long long* o = a.dbg + 256 * 8 + (tid ? 4 : 0);
        o[0] = tw; o[1] = tu; o[2] = tr;
    }
    // ------------------------------------------------------------ epilogue: state back to HBM
    if (!is_dec) {
#pragma unroll
        for (int r = 0; r < G; ++r) {
            if (r < g_act) {
                float4* dstf = reinterpret_cast<float4*>(a.fields + (size_t)(rep0 + r) * n_pad);
                uint32_t* dsts = reinterpret_cast<uint32_t*>(a.spins + (size_t)(rep0 + r) * n_pad);
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    dstf[k * kBulkThreads + tid] = make_float4(f[r][2 * k].x, f[r][2 * k].y,
                                                               f[r][2 * k + 1].x, f[r][2 * k + 1].y);
                    const uint32_t nib =
                        (sbits[r * W + k * (kBulkThreads / 8) + (tid >> 3)] >> ((tid & 7) * 4)) & 0xFu;
                    uint32_t bytes = 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        bytes |= (((nib >> e) & 1u) ? 0x01u : 0xFFu) << (8 * e);
                    dsts[k * kBulkThreads + tid] = bytes;
                }
            }
        }
    } else if (lane < g_act) {
        a.energy[rep0 + lane] = cur_e;
        if (a.track_best) a.best_energy[rep0 + lane] = best_e;
        a.accepted[rep0 + lane] += (unsigned long long)n_acc;
    }
}

template <int CPT, int G>
cudaError_t launch_t(const SweepDev& a, bool inject, int grid, cudaStream_t st) {
    const size_t smem = make_layout(a.n_pad, G, a.D).total;
    cudaError_t err;
    if (inject) {
        err = cudaFuncSetAttribute(sweep_kernel<CPT, G, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        sweep_kernel<CPT, G, true><<<grid, kSweepThreads, smem, st>>>(a);
    } else {
        err = cudaFuncSetAttribute(sweep_kernel<CPT, G, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        sweep_kernel<CPT, G, false><<<grid, kSweepThreads, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// replicas per block for each instantiated number of columns per thread (CPT = n_pad / 224)
__host__ __device__ constexpr int g_of_cpt(int cpt) {
    return cpt == 4 ? 32 : cpt == 8 ? 24 : cpt == 12 ? 16 : cpt == 16 ? 12 : cpt == 20 ? 10
         : cpt == 24 ? 8 : cpt == 28 ? 7 : cpt == 32 ? 6 : 0;
}

}  // namespace

int sweep_max_replicas_per_block(int n_pad) {
    if (n_pad % kColQuantum != 0) return 0;
    return g_of_cpt(n_pad / kBulkThreads);
}

size_t sweep_smem_bytes(int n_pad, int g, int D) { return make_layout(n_pad, g, D).total; }

cudaError_t launch_sweep(SweepDev a, bool inject, int grid, cudaStream_t st) {
    const int gt = sweep_max_replicas_per_block(a.n_pad);
    if (gt == 0 || a.G < 1 || a.G > gt) return cudaErrorInvalidValue;
    // deepest ring that fits in 227 KB of shared memory (and never deeper than a sweep)
    int D = kMaxStages;
    while (D > 2 && make_layout(a.n_pad, gt, D).total > 227 * 1024) --D;
    if (D > a.n) D = a.n;
    if (D < 2) return cudaErrorInvalidValue;  // the ring needs two stages (n >= 2)
    a.D = D;
    switch (a.n_pad / kBulkThreads) {
#define SG_LAUNCH(C) case C: return launch_t<C, g_of_cpt(C)>(a, inject, grid, st);
        SG_LAUNCH(4) SG_LAUNCH(8) SG_LAUNCH(12) SG_LAUNCH(16) SG_LAUNCH(20) SG_LAUNCH(24)
        SG_LAUNCH(28) SG_LAUNCH(32)
#undef SG_LAUNCH
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sg
