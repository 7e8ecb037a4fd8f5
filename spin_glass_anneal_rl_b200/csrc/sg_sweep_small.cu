// sg_sweep_small.cu -- K1-SMALL: the sweep for dense models whose couplings fit in shared memory
// (n <= 224: BASELINE cfg1, N = 100, and the 10..200-spin models the reference's RL environment
// anneals over and over, rl_integration/environment.py:318-336).
//
// Same algorithm, Philox counters, site orders and accept rules as the other kernels
// (SpinDynamics.sweep(), reference core/spin_dynamics.py:61-94, 131-191) -- but the parallel
// axis is turned round.  The big kernels put replicas on lanes and stream J rows past them; for a
// 100-spin model with 32 replicas that leaves one or two warps on the whole GPU.  Here ONE WARP
// OWNS ONE REPLICA: lane = field column, the replica's local fields are ceil(n/32) registers per
// lane, the spins one bit per column, J sits in shared memory once per CTA (8 replicas share it).
// An attempt is: shuffle the site's field and spin bit to every lane, decide (every lane computes
// the same decision), and on acceptance one FMA per register with a conflict-free row of J:
//       f_j <- fma(-2 s_i, J_ij, f_j)   for all j,
// the sequential algorithm's arithmetic exactly.  Sites, thresholds (or injected uniforms) are
// prepared 128 attempts at a time, four per lane, off the dependency chain.
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

namespace {

constexpr int kSmallWarps = 8;   // replicas per CTA

template <int NR, bool INJECT>
__global__ void __launch_bounds__(kSmallWarps * 32)
sweep_small_kernel(const SweepDev a, const int* __restrict__ sites_g) {
    extern __shared__ __align__(16) float Js[];   // [n][NR * 32], row i = column i of J, zero padded
    constexpr int ROW = NR * 32;
    const int n = a.n, n_pad = a.n_pad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // several models in one launch (a.rpm replicas each, couplings and fields stacked): a CTA
    // serves 8 replicas of ONE model
    const int rpm = a.rpm > 0 ? a.rpm : a.R;
    const int cpm = (rpm + kSmallWarps - 1) / kSmallWarps;
    const int model = blockIdx.x / cpm;
    const int rim = (blockIdx.x - model * cpm) * kSmallWarps + warp;
    const float* Jm = a.Jt + (size_t)model * n * n_pad;
    const float* hm = a.h + (size_t)model * n_pad;
    for (int idx = tid; idx < n * ROW; idx += kSmallWarps * 32) {
        const int i = idx / ROW, c = idx - i * ROW;
        Js[idx] = (c < n) ? Jm[(size_t)i * n_pad + c] : 0.0f;
    }
    __syncthreads();
    if (rim >= rpm) return;
    const int rep = model * rpm + rim;

    // this replica's state: column c = k * 32 + lane lives in f[k] / bit k of sb
    float f[NR], hv[NR];
    uint32_t sb = 0u;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const int c = k * 32 + lane;
        f[k] = (c < n) ? a.fields[(size_t)rep * n_pad + c] : 0.0f;
        hv[k] = (c < n) ? hm[c] : 0.0f;
        if (c < n && a.spins[(size_t)rep * n_pad + c] >= 0) sb |= 1u << k;
    }
    float cur_e = a.energy[rep];
    float best_e = a.track_best ? a.best_energy[rep] : 3.0e38f;
    unsigned int n_acc = 0;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));

#pragma unroll 1
    for (int s = 0; s < a.n_sweeps; ++s) {
        const int* tab = sites_g + (size_t)s * n;
        const unsigned long long sa = a.sweep_base + (unsigned long long)s;
        const double dT = a.temps[(long long)s * a.t_ss + (long long)rep * a.t_rs];
        const float Tm = (float)dT;
        const float* up = INJECT ? a.uniforms + ((size_t)rep * a.n_sweeps + s) * n : nullptr;
#pragma unroll 1
        for (int i0 = 0; i0 < n; i0 += 128) {
            // attempts i0 + 4 lane .. + 3: sites and thresholds (Philox counter = attempt / 4, as
            // everywhere) or injected uniforms
            int sq[4];
            float tq[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = i0 + 4 * lane + e;
                sq[e] = (i < n) ? tab[i] : 0;
                tq[e] = 0.0f;
                if (INJECT && i < n) tq[e] = up[i];
            }
            if (!INJECT && i0 + 4 * lane < n) {
                const uint4 x = philox4x32_10(
                    make_uint4((uint32_t)(rep + a.rep_base), (uint32_t)sa, (uint32_t)(sa >> 32), (uint32_t)((i0 >> 2) + lane)), key);
                const uint32_t vv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float u = u01(vv[e]);
                    tq[e] = (a.rule == 0) ? -__logf(u) * Tm : 0.5f * Tm * (__logf(u) - __logf(1.0f - u));
                }
            }
            const int nq = min(32, (n - i0 + 3) >> 2);
#pragma unroll 1
            for (int jq = 0; jq < nq; ++jq) {
                int site4[4];
                float th4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    site4[e] = __shfl_sync(0xFFFFFFFFu, sq[e], jq);
                    th4[e] = __shfl_sync(0xFFFFFFFFu, tq[e], jq);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (i0 + 4 * jq + e >= n) break;
                    const int site = site4[e];
                    const int k = site >> 5, src = site & 31;
                    float fsel = f[0];
#pragma unroll
                    for (int kk = 1; kk < NR; ++kk) fsel = (k == kk) ? f[kk] : fsel;
                    const float fv = __shfl_sync(0xFFFFFFFFu, fsel, src);
                    const bool upb = (__shfl_sync(0xFFFFFFFFu, sb, src) >> k) & 1u;
                    const float x = upb ? 2.0f * fv : -2.0f * fv;   // dE = 2 s f
                    bool flip;
                    if (!INJECT) {
                        flip = (a.rule == 0) ? (x < th4[e]) : ((fv > th4[e]) != upb);
                    } else {
                        const float u = th4[e];
                        if (a.rule == 0) {
                            // reference: dE <= 0 accepts without a draw; else u < exp(float(-dE/T))
                            flip = (x <= 0.0f) || (u < expf((float)(-(double)x / dT)));
                        } else {
                            const float arg = (a.rule == 1) ? (float)(-2.0 * (double)fv / dT)
                                                            : (float)(-2.0 * (1.0 / dT) * (double)fv);
                            const float p_up = 1.0f / (1.0f + expf(arg));
                            flip = ((u < p_up) != upb);
                        }
                    }
                    if (flip) {
                        const float d = upb ? -2.0f : 2.0f;
                        const float* row = Js + site * ROW + lane;
#pragma unroll
                        for (int kk = 0; kk < NR; ++kk) f[kk] = fmaf(d, row[kk * 32], f[kk]);
                        if (lane == src) sb ^= 1u << k;
                        ++n_acc;
                        if (a.site_de && lane == 0) a.site_de[(size_t)rep * n + site] += x;
                    }
                }
            }
        }
        // ---- end of sweep: E = -1/2 sum_c s_c (f_c + h_c), best tracking
        float part = 0.0f;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const float t = f[k] + hv[k];
            if (k * 32 + lane < n) part += ((sb >> k) & 1u) ? t : -t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
        cur_e = -0.5f * part;
        if (a.energy_trace && lane == 0) a.energy_trace[(size_t)s * a.R + rep] = cur_e;
        if (a.track_best && cur_e < best_e) {
            best_e = cur_e;
#pragma unroll
            for (int k = 0; k < NR; ++k) {
                const int c = k * 32 + lane;
                if (c < n) a.best_spins[(size_t)rep * n_pad + c] = ((sb >> k) & 1u) ? 1 : -1;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const int c = k * 32 + lane;
        if (c < n) {
            a.fields[(size_t)rep * n_pad + c] = f[k];
            a.spins[(size_t)rep * n_pad + c] = ((sb >> k) & 1u) ? 1 : -1;
        }
    }
    if (lane == 0) {
        a.energy[rep] = cur_e;
        if (a.track_best) a.best_energy[rep] = best_e;
        a.accepted[rep] += (unsigned long long)n_acc;
    }
}

// F = S J^T + h and E = -1/2 sum s (F + h) for stacked models, one warp per configuration
// (fp32, one FMA per coupling in column order: exact for integer couplings)
__global__ void __launch_bounds__(kSmallWarps * 32)
fields_small_kernel(const float* __restrict__ Jt, const float* __restrict__ h,
                    const int8_t* __restrict__ spins, float* __restrict__ fields,
                    float* __restrict__ energy, int n, int n_pad, int B, int rpm) {
    const int lane = threadIdx.x & 31;
    const int rep = blockIdx.x * kSmallWarps + (threadIdx.x >> 5);
    if (rep >= B) return;
    const int model = rep / rpm;
    const float* Jm = Jt + (size_t)model * n * n_pad;
    const float* hm = h + (size_t)model * n_pad;
    const int8_t* sp = spins + (size_t)rep * n_pad;
    float part = 0.0f;
    for (int c = lane; c < n; c += 32) {
        float acc = 0.0f;
        for (int j = 0; j < n; ++j) acc = fmaf(Jm[(size_t)j * n_pad + c], (float)sp[j], acc);
        const float f = acc + hm[c];
        if (fields) fields[(size_t)rep * n_pad + c] = f;
        const float t = f + hm[c];
        part += (sp[c] >= 0) ? t : -t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    if (lane == 0 && energy) energy[rep] = -0.5f * part;
}

// stacked models: Jt[m][i][c] = J[m][c][i] (c < n, else 0), h_pad[m][c] = h[m][c] (else 0)
__global__ void __launch_bounds__(256)
stack_models_kernel(const float* __restrict__ J, const float* __restrict__ h, int M, int n, int n_pad,
                    float* __restrict__ Jt, float* __restrict__ h_pad) {
    const size_t total = (size_t)M * n * n_pad;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const int c = (int)(idx % n_pad);
        const size_t mi = idx / n_pad;
        const int i = (int)(mi % n);
        const size_t m = mi / n;
        Jt[idx] = (c < n) ? J[(m * n + c) * n + i] : 0.0f;
        if (i == 0) h_pad[m * n_pad + c] = (c < n) ? h[m * n + c] : 0.0f;
    }
}

template <int NR>
cudaError_t launch_nr(const SweepDev& a, bool inject, const int* sites, cudaStream_t st) {
    const size_t smem = (size_t)a.n * NR * 32 * sizeof(float);
    const int rpm = a.rpm > 0 ? a.rpm : a.R;
    const int grid = (a.R / rpm) * ((rpm + kSmallWarps - 1) / kSmallWarps);
    cudaError_t e;
    if (inject) {
        e = cudaFuncSetAttribute(sweep_small_kernel<NR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sweep_small_kernel<NR, true><<<grid, kSmallWarps * 32, smem, st>>>(a, sites);
    } else {
        e = cudaFuncSetAttribute(sweep_small_kernel<NR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sweep_small_kernel<NR, false><<<grid, kSmallWarps * 32, smem, st>>>(a, sites);
    }
    return cudaGetLastError();
}

}  // namespace

bool sweep_small_supported(int n) { return n >= 1 && n <= 224; }

cudaError_t launch_stack_models(const float* J, const float* h, int M, int n, int n_pad, float* Jt,
                                float* h_pad, cudaStream_t st) {
    const size_t total = (size_t)M * n * n_pad;
    int grid = (int)((total + 255) / 256 < 4736 ? (total + 255) / 256 : 4736);
    stack_models_kernel<<<grid, 256, 0, st>>>(J, h, M, n, n_pad, Jt, h_pad);
    return cudaGetLastError();
}

cudaError_t launch_fields_small(const float* Jt, const float* h, const int8_t* spins, float* fields,
                                float* energy, int n, int n_pad, int B, int rpm, cudaStream_t st) {
    fields_small_kernel<<<(B + kSmallWarps - 1) / kSmallWarps, kSmallWarps * 32, 0, st>>>(
        Jt, h, spins, fields, energy, n, n_pad, B, rpm);
    return cudaGetLastError();
}

cudaError_t launch_sweep_small(const SweepDev& a, bool inject, const int* sites, cudaStream_t st) {
    if (!sweep_small_supported(a.n)) return cudaErrorInvalidValue;
    switch ((a.n + 31) / 32) {
        case 1: return launch_nr<1>(a, inject, sites, st);
        case 2: return launch_nr<2>(a, inject, sites, st);
        case 3: return launch_nr<3>(a, inject, sites, st);
        case 4: return launch_nr<4>(a, inject, sites, st);
        case 5: return launch_nr<5>(a, inject, sites, st);
        case 6: return launch_nr<6>(a, inject, sites, st);
        default: return launch_nr<7>(a, inject, sites, st);
    }
}

}  // namespace sg
