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
//     copies (cp.async.bulk + mbarrier) through a D-stage ring.
//   * The 8 warps keep the local fields f[r][j] = h_j + sum_i J_ji s_ri of all
//     G replicas RESIDENT IN REGISTERS: thread t owns columns {1024k + 4t + e}.
//     An accepted flip of spin i in replica r is the rank-1 update
//     f[r][:] += -2 s_ri * Jt[i][:]  (incremental field update, one FFMA per column).
//   * The accept decision for the NEXT attempt is made while the current one is
//     being applied: the warp that owns the next site's column transposes the G
//     field values of that column through shared memory (lane = replica), adds the
//     pending update of that single column itself (the same FFMA the owner thread
//     executes a moment later, so the value is bit-identical), compares against the
//     pre-computed Metropolis/Glauber threshold and publishes a (flip, sign) mask
//     that every warp reads after the per-attempt barrier.
//     (A dedicated 9th decision warp was tried first: 9 warps put 3 warps on one
//     SM sub-partition and cap every thread at 168 registers, i.e. G = 8.)
//   * Thresholds: accept <=> dE < -T ln(u).  The u's come from Philox4x32-10 keyed
//     on (replica, absolute sweep, attempt) and are produced 32 attempts ahead by
//     the bulk warps, off the critical path.  In injected mode the caller supplies
//     the uniforms and the reference's own comparison u < exp(-dE/T) is evaluated.
//   * Spins live as bit planes in shared memory (deciding lane r owns plane r).
//   * After every sweep the energy of each replica is reduced from the resident
//     fields, E = -1/2 sum_j s_j (f_j + h_j), and the best configuration is kept.
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

struct SmemLayout {
    size_t jring, sites, sbits, theta, red, xfer, acc, pub, flags, mbar, total;
};

__host__ __device__ inline SmemLayout make_layout(int n_pad, int G, int D) {
    SmemLayout L;
    size_t off = 0;
    L.jring = off; off += (size_t)D * n_pad * sizeof(float);
    L.sites = off; off += 2 * (size_t)n_pad * sizeof(uint16_t);
    L.sbits = off; off += (size_t)G * (n_pad / 32) * sizeof(uint32_t);
    L.theta = off; off += 2 * kSB * 32 * sizeof(float);
    L.red = off;   off += 8 * 32 * sizeof(float);
    L.xfer = off;  off += 32 * sizeof(float);
    L.acc = off;   off += 32 * sizeof(uint32_t);
    L.pub = off;   off += 4 * sizeof(uint2);
    L.flags = off; off += 4 * sizeof(uint32_t);
    off = (off + 15) & ~(size_t)15;
    L.mbar = off;  off += kMaxStages * sizeof(uint64_t);
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
    float* xfer = reinterpret_cast<float*>(smem + L.xfer);
    uint32_t* acc_s = reinterpret_cast<uint32_t*>(smem + L.acc);
    uint2* pub = reinterpret_cast<uint2*>(smem + L.pub);
    uint32_t* flags = reinterpret_cast<uint32_t*>(smem + L.flags);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.mbar);

    const int rep0 = blockIdx.x * a.G;
    const int g_act = min(a.G, a.R - rep0);
    const int n_sweeps = a.n_sweeps;
    const long long total = (long long)n_sweeps * n;
    const int nb = (n + kSB - 1) / kSB;  // threshold batches per sweep
    const uint32_t row_bytes = (uint32_t)n_pad * 4u;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));

    // ------------------------------------------------------------ helpers
    auto gen_sites = [&](int s) {  // site table of launch-local sweep s
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

    auto gen_theta = [&](int s, int bi) {  // thresholds of batch bi of sweep s
        if (INJECT || s >= n_sweeps) return;
        const int r = tid & 31, q = tid >> 5;
        const int i0 = bi * kSB + q * 4;
        if (r >= g_act || i0 >= n) return;
        const int rep = rep0 + r;
        const unsigned long long sa = a.sweep_base + (unsigned long long)s;
        const float T = (float)a.temps[(long long)s * a.t_ss + (long long)rep * a.t_rs];
        const uint4 x = philox4x32_10(
            make_uint4((uint32_t)rep, (uint32_t)sa, (uint32_t)(sa >> 32), (uint32_t)(i0 >> 2)), key);
        const uint32_t v[4] = {x.x, x.y, x.z, x.w};
        float* dst = theta + (size_t)((s * nb + bi) & 1) * (kSB * 32) + (q * 4) * 32 + r;
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
        for (int d = 0; d < D; ++d) mbar_init(&full[d], 1);
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

    // prefetch cursor (only thread 0 issues TMA)
    int pf_s = 0, pf_i = 0;
    auto issue_row = [&](int stage) {
        const int site = sites_s[(size_t)(pf_s & 1) * n_pad + pf_i];
        mbar_arrive_expect_tx(&full[stage], row_bytes);
        bulk_g2s(Jring + (size_t)stage * n_pad, a.Jt + (size_t)site * n_pad, row_bytes,
                 &full[stage]);
        if (++pf_i == n) { pf_i = 0; ++pf_s; }
    };
    if (tid == 0) {
        for (int d = 0; d < D && d < total; ++d) issue_row(d);
    }

    // Decision for attempt (s, i) at `site`, executed by the warp that owns the site's
    // column.  (P, Jv) is the flip that is still being applied: its contribution to this one
    // column is added here with the same FFMA the owner thread executes in the bulk update.
    auto decide = [&](int s, int i, int site, uint2 P, float Jv, int slot) {
        const int ot = (site >> 2) & (kSweepThreads - 1);
        if ((ot >> 5) != warp) return;
        if ((ot & 31) == lane) publish_column<CPT, G>(f, ((site >> 10) << 2) | (site & 3), xfer);
        __syncwarp();
        const float v = xfer[lane];
        bool flip = false, s_up = false;
        if (lane < g_act) {
            const float d = ((P.x >> lane) & 1u) ? (((P.y >> lane) & 1u) ? -2.0f : 2.0f) : 0.0f;
            const float v2 = fmaf(d, Jv, v);  // local field of `site` after the pending flip
            uint32_t* wp = &sbits[lane * W + (site >> 5)];
            const uint32_t w = *wp;
            s_up = (w >> (site & 31)) & 1u;
            if (!INJECT) {
                const float th =
                    theta[(size_t)((s * nb + (i >> 5)) & 1) * (kSB * 32) + (i & 31) * 32 + lane];
                if (a.rule == 0) {
                    const float x = s_up ? 2.0f * v2 : -2.0f * v2;  // dE = 2 s f
                    flip = x < th;
                } else {
                    flip = ((v2 > th) != s_up);
                }
            } else {
                const float u = a.uniforms[((size_t)(rep0 + lane) * n_sweeps + s) * n + i];
                const double T_d = a.temps[(long long)s * a.t_ss + (long long)(rep0 + lane) * a.t_rs];
                if (a.rule == 0) {
                    const float x = s_up ? 2.0f * v2 : -2.0f * v2;
                    // reference: dE <= 0 accepts without a draw; else u < exp(float(-dE/T))
                    flip = (x <= 0.0f) || (u < expf((float)(-(double)x / T_d)));
                } else {
                    const float arg = (a.rule == 1) ? (float)(-2.0 * (double)v2 / T_d)
                                                    : (float)(-2.0 * (1.0 / T_d) * (double)v2);
                    const float p_up = 1.0f / (1.0f + expf(arg));
                    flip = ((u < p_up) != s_up);
                }
            }
            if (flip) {
                *wp = w ^ (1u << (site & 31));
                acc_s[lane] += 1u;
            }
        }
        const uint32_t am = __ballot_sync(0xFFFFFFFFu, flip);
        const uint32_t sm = __ballot_sync(0xFFFFFFFFu, s_up);
        if (lane == 0) pub[slot] = make_uint2(am, sm);
        __syncwarp();  // xfer is reused by the next decision of this warp
    };

    // ------------------------------------------------------------ sweeps
    long long g = 0;  // launch-local attempt counter
    int stage = 0;
    uint32_t parity = 0;
    for (int s = 0; s < n_sweeps; ++s) {
        const uint16_t* tab = sites_s + (size_t)(s & 1) * n_pad;
        if (s >= 1 && s + 1 < n_sweeps) gen_sites(s + 1);

        // first attempt of the sweep: nothing pending
        decide(s, 0, tab[0], make_uint2(0u, 0u), 0.0f, (int)(g & 3));
        __syncthreads();

        for (int i = 0; i < n; ++i) {
            const uint2 P = pub[g & 3];
            mbar_wait(&full[stage], parity);
            const float* Jrow = Jring + (size_t)stage * n_pad;

            if (i + 1 < n) {
                const int sn = tab[i + 1];
                decide(s, i + 1, sn, P, Jrow[sn], (int)((g + 1) & 3));
            }
            if ((i & (kSB - 1)) == 0) {  // thresholds for the batch after this one
                if (i + kSB < n) gen_theta(s, (i >> 5) + 1);
                else gen_theta(s + 1, 0);
            }
            if (P.x != 0u) {
                const float4* Jr4 = reinterpret_cast<const float4*>(Jrow);
                float4 jv[KCH];
#pragma unroll
                for (int k = 0; k < KCH; ++k) jv[k] = Jr4[k * kSweepThreads + tid];
#pragma unroll
                for (int r = 0; r < G; ++r) {
                    if (P.x & (1u << r)) {
                        const float d = (P.y & (1u << r)) ? -2.0f : 2.0f;
#pragma unroll
                        for (int k = 0; k < KCH; ++k) {
                            f[r][4 * k + 0] = fmaf(d, jv[k].x, f[r][4 * k + 0]);
                            f[r][4 * k + 1] = fmaf(d, jv[k].y, f[r][4 * k + 1]);
                            f[r][4 * k + 2] = fmaf(d, jv[k].z, f[r][4 * k + 2]);
                            f[r][4 * k + 3] = fmaf(d, jv[k].w, f[r][4 * k + 3]);
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0 && g + D < total) issue_row(stage);
            ++g;
            if (++stage == D) { stage = 0; parity ^= 1u; }
        }

        // ---- end of sweep: energies from the resident fields, best tracking
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
