// sg_sweep.cu -- K1: the replica-batched Monte Carlo sweep for sm_100a.
//
// Replaces SpinDynamics.sweep() / _metropolis_update / _glauber_update /
// _heat_bath_update (reference core/spin_dynamics.py:61-94, 131-191) and the
// never-launched metropolis_update_kernel (annealing/cuda_kernels.py:8-50).
//
// Mapping (one thread block per SM, 256 threads = 8 warps, up to 255 registers each):
//   * A block owns G replicas and visits the sites of a sweep in ONE order shared
//     by its replicas, so row Jt[site][:] is fetched once per attempt and serves
//     all G replicas.  Rows are streamed L2/HBM -> shared memory by TMA bulk
//     copies (cp.async.bulk + mbarrier) through a D-stage ring; a stage is
//     recycled through an "empty" mbarrier on which every warp arrives.
//   * The 8 warps keep the local fields f[r][j] = h_j + sum_i J_ji s_ri of all
//     G replicas RESIDENT IN REGISTERS: thread t owns columns {1024k + 4t + e}.
//     An accepted flip of spin i in replica r is the rank-1 update
//     f[r][:] += -2 s_ri * Jt[i][:]  (incremental field update, one FFMA per column).
//   * No block-wide barrier inside a sweep.  The accept decisions form a chain
//     through a ring of published (flip mask, sign mask, tag) words in shared
//     memory: the warp that owns the column of attempt j+1 prepares everything that
//     does not depend on attempt j (its G raw field values transposed lane=replica,
//     the coupling J[site_j][site_j+1], spin bits, thresholds), polls the ring for
//     decision j, adds the pending contribution of that one flip with the same FFMA
//     it will execute in its own bulk update a moment later (bit-identical value),
//     compares, and publishes decision j+1.  Every warp then applies decision j to
//     its registers at its own pace; warps may drift apart by up to D attempts.
//     (v1 used a __syncthreads per attempt and cost ~1500 clk/attempt; a dedicated
//     9th decision warp caps every thread at 168 registers, i.e. G = 8.)
//   * Thresholds: accept <=> dE < -T ln(u).  The u's come from Philox4x32-10 keyed
//     on (replica, absolute sweep, attempt) and are produced one 32-attempt batch
//     ahead, off the decision chain.  In injected mode the caller supplies the
//     uniforms and the reference's own comparison u < exp(-dE/T) is evaluated.
//   * Spins live as bit planes in shared memory (deciding lane r owns plane r).
//   * After every sweep (block barrier) the energy of each replica is reduced from
//     the resident fields, E = -1/2 sum_j s_j (f_j + h_j), and the best
//     configuration is kept.
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

namespace {

constexpr int kSB = 32;                      // attempts per threshold batch
constexpr int kMaxStages = 8;

template <int C, int CPT, int G>
__device__ __forceinline__ void publish_one(const float (&f)[G][CPT], float* xfer) {
    if constexpr (C < CPT) {
#pragma unroll
        for (int r = 0; r < G; ++r) xfer[r] = f[r][C];
    }
}

// the owner thread of a site copies its G field values of local column c to shared memory
// (a switch, so that every register index is a compile-time constant)
template <int CPT, int G>
__device__ __forceinline__ void publish_column(const float (&f)[G][CPT], int c, float* xfer) {
#define SG_CASE(C) case C: publish_one<C, CPT, G>(f, xfer); break;
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

constexpr int kRing = 16;   // published-decision ring (must exceed kMaxStages + 1)
constexpr int kThetaBufs = 3;

struct SmemLayout {
    size_t jring, sites, jvtab, sbits, theta, red, xfer, acc, pub, flags, mbar, total;
};

__host__ __device__ inline SmemLayout make_layout(int n_pad, int G, int D) {
    SmemLayout L;
    size_t off = 0;
    L.jring = off; off += (size_t)D * n_pad * sizeof(float);
    L.sites = off; off += 2 * (size_t)n_pad * sizeof(uint16_t);
    L.jvtab = off; off += (size_t)n_pad * sizeof(float);
    L.sbits = off; off += (size_t)G * (n_pad / 32) * sizeof(uint32_t);
    L.theta = off; off += (size_t)kThetaBufs * kSB * 32 * sizeof(float);
    L.red = off;   off += 8 * 32 * sizeof(float);
    L.xfer = off;  off += 8 * 32 * sizeof(float);
    L.acc = off;   off += 32 * sizeof(uint32_t);
    L.pub = off;   off += kRing * (32 * sizeof(float) + sizeof(uint32_t));
    L.flags = off; off += 4 * sizeof(uint32_t);
    off = (off + 15) & ~(size_t)15;
    L.mbar = off;  off += (2 * kMaxStages + kRing) * sizeof(uint64_t);
    L.total = off;
    return L;
}


template <int CPT, int G, bool INJECT>
__global__ void __launch_bounds__(kSweepThreads, 1) sweep_kernel(const SweepDev a) {
    constexpr int KCH = CPT / 4;
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = a.n, n_pad = a.n_pad, D = a.D;
    const int W = n_pad >> 5;  // spin-bit words per replica
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const SmemLayout L = make_layout(n_pad, G, D);
    float* Jring = reinterpret_cast<float*>(smem + L.jring);
    uint16_t* sites_s = reinterpret_cast<uint16_t*>(smem + L.sites);
    uint32_t* sbits = reinterpret_cast<uint32_t*>(smem + L.sbits);
    float* theta = reinterpret_cast<float*>(smem + L.theta);
    float* red = reinterpret_cast<float*>(smem + L.red);
    float* xfer = reinterpret_cast<float*>(smem + L.xfer) + warp * 32;
    uint32_t* acc_s = reinterpret_cast<uint32_t*>(smem + L.acc);
    float* dpub = reinterpret_cast<float*>(smem + L.pub);  // [kRing][32] field deltas (-2 s) of a decision
    uint32_t* ampub = reinterpret_cast<uint32_t*>(dpub + kRing * 32);  // [kRing] flip masks
    float* jvtab = reinterpret_cast<float*>(smem + L.jvtab);
    uint32_t* flags = reinterpret_cast<uint32_t*>(smem + L.flags);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.mbar);
    uint64_t* empty = full + kMaxStages;
    uint64_t* pbar = empty + kMaxStages;  // one per slot of the decision ring

    const int rep0 = blockIdx.x * a.G;
    const int g_act = min(a.G, a.R - rep0);
    const int n_sweeps = a.n_sweeps;
    const long long total = (long long)n_sweeps * n;
    const int nb = (n + kSB - 1) / kSB;  // threshold batches per sweep
    const uint32_t row_bytes = (uint32_t)n_pad * 4u;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));

    // ------------------------------------------------------------ helpers
    auto gen_sites = [&](int s) {  // site table of launch-local sweep s (all threads)
        uint16_t* tab = sites_s + (size_t)(s & 1) * n_pad;
        if (a.site_mode == 0) {
            for (int i = tid; i < n; i += kSweepThreads) tab[i] = (uint16_t)i;
        } else if (a.site_mode == 1) {
            const unsigned long long sa = a.sweep_base + (unsigned long long)s;
            for (int q = tid; q * 4 < n; q += kSweepThreads) {
                const uint4 x = philox4x32_10(
                    make_uint4(kSiteStreamTag, (uint32_t)sa, (uint32_t)(sa >> 32), (uint32_t)q), key);
                const uint32_t v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (q * 4 + e < n) tab[q * 4 + e] = (uint16_t)(v[e] % (uint32_t)n);
            }
        } else {
            const int* src = a.sites + (long long)blockIdx.x * a.s_bs + (long long)s * a.s_ss;
            for (int i = tid; i < n; i += kSweepThreads) tab[i] = (uint16_t)src[i];
        }
    };

    // thresholds of batch bi of sweep s: warp q produces attempts 4q..4q+3 for lane = replica
    auto gen_theta = [&](int s, int bi) {
        if (INJECT || s >= n_sweeps) return;
        const int r = lane, q = warp;
        const int i0 = bi * kSB + q * 4;
        if (r >= g_act || i0 >= n) return;
        const int rep = rep0 + r;
        const unsigned long long sa = a.sweep_base + (unsigned long long)s;
        const float T = (float)a.temps[(long long)s * a.t_ss + (long long)rep * a.t_rs];
        const uint4 x = philox4x32_10(
            make_uint4((uint32_t)rep, (uint32_t)sa, (uint32_t)(sa >> 32), (uint32_t)(i0 >> 2)), key);
        const uint32_t v[4] = {x.x, x.y, x.z, x.w};
        float* dst = theta + (size_t)((s * nb + bi) % kThetaBufs) * (kSB * 32) + (q * 4) * 32 + r;
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
            mbar_init(&empty[d], kSweepThreads / 32);
        }
        for (int d = 0; d < kRing; ++d) mbar_init(&pbar[d], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    if (tid < 32) acc_s[tid] = 0u;
    gen_sites(0);
    if (n_sweeps > 1) gen_sites(1);
    gen_theta(0, 0);
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

    // resident local fields
    float f[G][CPT];
#pragma unroll
    for (int r = 0; r < G; ++r) {
        if (r < g_act) {
            const float4* src =
                reinterpret_cast<const float4*>(a.fields + (size_t)(rep0 + r) * n_pad);
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                const float4 v = src[k * kSweepThreads + tid];
                f[r][4 * k + 0] = v.x; f[r][4 * k + 1] = v.y;
                f[r][4 * k + 2] = v.z; f[r][4 * k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < CPT; ++c) f[r][c] = 0.0f;
        }
    }

    // per-replica energies live in the lanes of warp 0
    float best_e = 3.0e38f, cur_e = 0.0f;
    if (warp == 0 && lane < g_act) {
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

    // ---- decision state of this warp (valid between prepare and finish)
    float dv = 0.0f, dJv = 0.0f, dth = 0.0f, du = 0.0f;
    double dT = 1.0;
    uint32_t dw = 0;
    uint32_t* dwp = nullptr;
    int dsite = 0;
    bool dup = false;

    // everything about attempt (s, i) at `site` that does not depend on the previous decision
    auto prepare = [&](int s, int i, int site) {
        const int ot = (site >> 2) & (kSweepThreads - 1);
        if ((ot & 31) == lane) publish_column<CPT, G>(f, ((site >> 10) << 2) | (site & 3), xfer);
        __syncwarp();
        dsite = site;
        dJv = (i > 0) ? jvtab[i] : 0.0f;  // J[site_i][site_{i-1}]: what the pending flip adds
        if (lane < g_act) {
            dv = xfer[lane];
            dwp = &sbits[lane * W + (site >> 5)];
            if (!INJECT) {
                dth = theta[(size_t)((s * nb + (i >> 5)) % kThetaBufs) * (kSB * 32) +
                            (i & 31) * 32 + lane];
            } else {
                du = a.uniforms[((size_t)(rep0 + lane) * n_sweeps + s) * n + i];
                dT = a.temps[(long long)s * a.t_ss + (long long)(rep0 + lane) * a.t_rs];
            }
        }
        __syncwarp();  // xfer may be rewritten by this warp's next prepare
    };

    // finish the prepared decision given the pending delta of the previous attempt (d_prev =
    // -2 s_old if that attempt flipped this lane's replica, else 0) and publish it as attempt g1
    auto finish = [&](float d_prev, long long g1) {
        bool flip = false;
        dup = false;
        if (lane < g_act) {
            const float v2 = fmaf(d_prev, dJv, dv);  // local field of the site after the pending flip
            dw = *dwp;                               // spin bits (the previous decider may have flipped)
            dup = (dw >> (dsite & 31)) & 1u;
            if (!INJECT) {
                if (a.rule == 0) {
                    const float x = dup ? 2.0f * v2 : -2.0f * v2;  // dE = 2 s f
                    flip = x < dth;
                } else {
                    flip = ((v2 > dth) != dup);
                }
            } else {
                if (a.rule == 0) {
                    const float x = dup ? 2.0f * v2 : -2.0f * v2;
                    // reference: dE <= 0 accepts without a draw; else u < exp(float(-dE/T))
                    flip = (x <= 0.0f) || (du < expf((float)(-(double)x / dT)));
                } else {
                    const float arg = (a.rule == 1) ? (float)(-2.0 * (double)v2 / dT)
                                                    : (float)(-2.0 * (1.0 / dT) * (double)v2);
                    const float p_up = 1.0f / (1.0f + expf(arg));
                    flip = ((du < p_up) != dup);
                }
            }
            if (flip) {
                *dwp = dw ^ (1u << (dsite & 31));
                acc_s[lane] += 1u;
            }
        }
        const int slot1 = (int)(g1 & (kRing - 1));
        dpub[slot1 * 32 + lane] = flip ? (dup ? -2.0f : 2.0f) : 0.0f;
        const uint32_t am = __ballot_sync(0xFFFFFFFFu, flip);
        __syncwarp();  // every lane's bit-plane / counter / delta store precedes the release below
        if (lane == 0) {
            ampub[slot1] = am;
            mbar_arrive(&pbar[slot1]);  // release: wakes the warps waiting for attempt g1
        }
    };

    // ------------------------------------------------------------ sweeps
    long long g = 0;  // launch-local attempt counter
    int stage = 0;
    uint32_t parity = 0;
#pragma unroll 1
    for (int s = 0; s < n_sweeps; ++s) {
        const uint16_t* tab = sites_s + (size_t)(s & 1) * n_pad;
        if (s >= 1 && s + 1 < n_sweeps) gen_sites(s + 1);
        // couplings between consecutive sites of this sweep (what a pending flip adds to the
        // next site's field): gathered once per sweep so the decision chain never waits on TMA
        for (int i = tid + 1; i < n; i += kSweepThreads)
            jvtab[i] = a.Jt[(size_t)tab[i - 1] * n_pad + tab[i]];
        __syncthreads();

        // first attempt of the sweep: nothing pending.  A decision is only published once the
        // row of its site has landed, so the other warps never wait on the TMA barrier.
        {
            const int site0 = tab[0];
            if ((((site0 >> 2) & (kSweepThreads - 1)) >> 5) == warp) {
                mbar_wait(&full[stage], parity);
                prepare(s, 0, site0);
                finish(0.0f, g);
            }
        }

#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const int slot = (int)(g & (kRing - 1));
            bool own = false;
            if (i + 1 < n) {
                const int sn = tab[i + 1];
                own = ((((sn >> 2) & (kSweepThreads - 1)) >> 5) == warp);
                if (own) {
                    const int nst = (stage + 1 == D) ? 0 : stage + 1;
                    mbar_wait(&full[nst], (nst == 0) ? (parity ^ 1u) : parity);  // row of g+1 landed
                    prepare(s, i + 1, sn);
                }
            }
            // decision of attempt g (hardware-suspended wait on the slot's mbarrier)
            mbar_wait(&pbar[slot], (uint32_t)(g >> 4) & 1u);
            const uint32_t am = ampub[slot];
            if (own) finish(dpub[slot * 32 + lane], g + 1);

            if ((i & (kSB - 1)) == 0) {  // thresholds for the batch after this one
                if (i + kSB < n) gen_theta(s, (i >> 5) + 1);
                else gen_theta(s + 1, 0);
            }
            if (am != 0u) {
                const float4* Jr4 = reinterpret_cast<const float4*>(Jring + (size_t)stage * n_pad);
                const float4* d4 = reinterpret_cast<const float4*>(dpub + slot * 32);
                float4 jv[KCH];
#pragma unroll
                for (int k = 0; k < KCH; ++k) jv[k] = Jr4[k * kSweepThreads + tid];
                // groups of 4 replicas: one uniform branch per group, FMAs with delta 0 for the
                // replicas of the group that did not flip (no per-replica branches)
#pragma unroll
                for (int q = 0; q < (G + 3) / 4; ++q) {
                    if ((am >> (4 * q)) & 0xFu) {
                        const float4 dq = d4[q];
                        const float dd[4] = {dq.x, dq.y, dq.z, dq.w};
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            const int r = 4 * q + rr;
                            if (r < G) {
#pragma unroll
                                for (int k = 0; k < KCH; ++k) {
                                    f[r][4 * k + 0] = fmaf(dd[rr], jv[k].x, f[r][4 * k + 0]);
                                    f[r][4 * k + 1] = fmaf(dd[rr], jv[k].y, f[r][4 * k + 1]);
                                    f[r][4 * k + 2] = fmaf(dd[rr], jv[k].z, f[r][4 * k + 2]);
                                    f[r][4 * k + 3] = fmaf(dd[rr], jv[k].w, f[r][4 * k + 3]);
                                }
                            }
                        }
                    }
                }
            }
            // release the stage.  Warp ((g-1) mod 8) re-arms the stage of the PREVIOUS attempt
            // with the row of attempt g-1+D: one attempt late, so that it rarely has to wait
            // for the slowest warp (rows are therefore D-1 attempts ahead).
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty[stage]);
                if (g >= 1 && (int)((g - 1) & 7) == warp && g - 1 + D < total) {
                    const int pst = (stage == 0) ? D - 1 : stage - 1;
                    const uint32_t ppar = (stage == 0) ? (parity ^ 1u) : parity;
                    mbar_wait(&empty[pst], ppar);  // all 8 warps are done with attempt g-1
                    mbar_wait(&full[pst], ppar);   // and its row landed (phase order of `full`)
                    int pi = i - 1 + D, ps = s;
                    if (pi >= n) { pi -= n; ++ps; }
                    issue_row(pst, ps, pi);
                }
            }
            ++g;
            if (++stage == D) { stage = 0; parity ^= 1u; }
        }

        // ---- end of sweep: energies from the resident fields, best tracking
        __syncthreads();
        {
            const float4* h4 = reinterpret_cast<const float4*>(a.h);
            float hv[CPT];
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                const float4 v = h4[k * kSweepThreads + tid];
                hv[4 * k + 0] = v.x; hv[4 * k + 1] = v.y; hv[4 * k + 2] = v.z; hv[4 * k + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < G; ++r) {
                float part = 0.0f;
#pragma unroll
                for (int k = 0; k < KCH; ++k) {
                    const uint32_t nib = (sbits[r * W + k * 32 + (tid >> 3)] >> ((tid & 7) * 4)) & 0xFu;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float t = f[r][4 * k + e] + hv[4 * k + e];
                        part += ((nib >> e) & 1u) ? t : -t;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
                if (lane == 0) red[warp * 32 + r] = part;
            }
        }
        __syncthreads();
        if (warp == 0) {
            bool improved = false;
            if (lane < g_act) {
                float acc = 0.0f;
#pragma unroll
                for (int w = 0; w < 8; ++w) acc += red[w * 32 + lane];
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
        if (im != 0u) {
#pragma unroll
            for (int r = 0; r < G; ++r) {
                if (im & (1u << r)) {
                    uint32_t* dst =
                        reinterpret_cast<uint32_t*>(a.best_spins + (size_t)(rep0 + r) * n_pad);
#pragma unroll
                    for (int k = 0; k < KCH; ++k) {
                        const uint32_t nib =
                            (sbits[r * W + k * 32 + (tid >> 3)] >> ((tid & 7) * 4)) & 0xFu;
                        uint32_t bytes = 0;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            bytes |= (((nib >> e) & 1u) ? 0x01u : 0xFFu) << (8 * e);
                        dst[k * kSweepThreads + tid] = bytes;
                    }
                }
            }
        }
        __syncthreads();  // bit planes / flags are modified again by the next sweep
    }

    // ------------------------------------------------------------ epilogue: state back to HBM
#pragma unroll
    for (int r = 0; r < G; ++r) {
        if (r < g_act) {
            float4* dstf = reinterpret_cast<float4*>(a.fields + (size_t)(rep0 + r) * n_pad);
            uint32_t* dsts = reinterpret_cast<uint32_t*>(a.spins + (size_t)(rep0 + r) * n_pad);
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                dstf[k * kSweepThreads + tid] = make_float4(f[r][4 * k + 0], f[r][4 * k + 1],
                                                            f[r][4 * k + 2], f[r][4 * k + 3]);
                const uint32_t nib = (sbits[r * W + k * 32 + (tid >> 3)] >> ((tid & 7) * 4)) & 0xFu;
                uint32_t bytes = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) bytes |= (((nib >> e) & 1u) ? 0x01u : 0xFFu) << (8 * e);
                dsts[k * kSweepThreads + tid] = bytes;
            }
        }
    }
    if (warp == 0 && lane < g_act) {
        a.energy[rep0 + lane] = cur_e;
        if (a.track_best) a.best_energy[rep0 + lane] = best_e;
        a.accepted[rep0 + lane] += (unsigned long long)acc_s[lane];
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

// (columns per thread, replicas per block) per padded size
constexpr int kG4 = 32, kG8 = 24, kG16 = 12, kG32 = 6;

int g_template(int n_pad) {
    switch (n_pad / kSweepThreads) {
        case 4: return kG4;
        case 8: return kG8;
        case 16: return kG16;
        case 32: return kG32;
        default: return 0;
    }
}

}  // namespace

int sweep_max_replicas_per_block(int n_pad) { return g_template(n_pad); }

size_t sweep_smem_bytes(int n_pad, int g, int D) { return make_layout(n_pad, g, D).total; }

cudaError_t launch_sweep(SweepDev a, bool inject, int grid, cudaStream_t st) {
    const int gt = g_template(a.n_pad);
    if (gt == 0 || a.G < 1 || a.G > gt) return cudaErrorInvalidValue;
    // deepest ring that fits in 227 KB of shared memory (and never deeper than a sweep)
    int D = kMaxStages;
    while (D > 1 && make_layout(a.n_pad, gt, D).total > 227 * 1024) --D;
    if (D > a.n) D = a.n;
    if (D < 2) return cudaErrorInvalidValue;  // the ring needs two stages (n >= 2)
    a.D = D;
    switch (a.n_pad / kSweepThreads) {
        case 4: return launch_t<4, kG4>(a, inject, grid, st);
        case 8: return launch_t<8, kG8>(a, inject, grid, st);
        case 16: return launch_t<16, kG16>(a, inject, grid, st);
        case 32: return launch_t<32, kG32>(a, inject, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sg
