// sg_sweep_groups.cu -- K1-GRP: the sweep for block-clique couplings (BASELINE cfg5).
//
// The one-hot / cardinality penalties the reference's constraint encoder produces
// (core/constraints.py:126-158 -> EqualityConstraint, :73-92) couple every pair of spins of a
// group with the SAME value: J_ij = c_g for i != j in group g, 0 otherwise.  SimpleScheduler
// (problems/simple_scheduler.py:67-127: 500 tasks x 100 agents, one group per task) is exactly
// that.  Then the local field needs no row of J at all:
//       f_i = h_i + c_g (S_g - s_i),      S_g = sum of the spins of group g,
// and a flip of spin i only changes S_g.  Same accept rules, Philox counters and site orders as
// the other kernels (SpinDynamics.sweep(), reference core/spin_dynamics.py:61-94, 131-191), so on
// integer data the trajectories are identical to the sparse kernel's.
//
// One warp = 32 replicas (lane = replica) with its whole state in shared memory: bit b of
// word[site] is the spin of replica b, sum[g][b] an int16.  Per attempt: two shared-memory reads,
// one compare, at most two writes; h_i, the group id and the Philox thresholds are prefetched.
// No field array exists, so nothing drifts: f is recomputed from the integer S_g every time.
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

namespace {

template <bool INJECT>
__global__ void __launch_bounds__(32, 1)
sweep_groups_kernel(const GrpDev m, const SweepDev a, const int* __restrict__ sites_g) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, NGp = m.n_groups;
    uint32_t* word = reinterpret_cast<uint32_t*>(smem);                 // [n]
    short* sum = reinterpret_cast<short*>(smem + (size_t)n * 4);        // [n_groups][32]
    const int lane = threadIdx.x;
    const int rep = blockIdx.x * 32 + lane;
    const bool active = rep < a.R;
    const int n_sweeps = a.n_sweeps;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    uint32_t* gw = m.words + (size_t)blockIdx.x * n;

    // state: spin words from HBM, group sums recomputed
    for (int i = lane; i < n; i += 32) word[i] = gw[i];
    for (int g = 0; g < NGp; ++g) sum[g * 32 + lane] = 0;
    __syncwarp();
    for (int i = 0; i < n; ++i) {
        const int g = m.group_of[i];
        sum[g * 32 + lane] += ((word[i] >> lane) & 1u) ? 1 : -1;
    }
    __syncwarp();

    float cur_e = active ? a.energy[rep] : 0.0f;
    float best_e = (active && a.track_best) ? a.best_energy[rep] : 3.0e38f;
    unsigned int n_acc = 0;

#pragma unroll 1
    for (int s = 0; s < n_sweeps; ++s) {
        const int* tab = sites_g + (size_t)s * n;
        const unsigned long long sa = a.sweep_base + (unsigned long long)s;
        const double dT = active ? a.temps[(long long)s * a.t_ss + (long long)rep * a.t_rs] : 1.0;
        const float Tm = (float)dT;
        const float* up = INJECT ? a.uniforms + ((size_t)rep * n_sweeps + s) * n : nullptr;
        float th4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        // state-independent per-attempt data (site, group, coupling, field) is loaded 32 attempts at
        // a time, one attempt per lane, a batch ahead, and broadcast with shuffles
        int b_site = 0, b_g = 0;
        float b_h = 0.0f, b_c = 0.0f;
        auto load_batch = [&](int i0, int& ls, int& lg, float& lh, float& lc) {
            const int ii = i0 + lane;
            ls = 0; lg = 0; lh = 0.0f; lc = 0.0f;
            if (ii < n) {
                ls = tab[ii];
                lg = m.group_of[ls];
                lh = m.h[ls];
                lc = m.coupling[lg];
            }
        };
        int n_site, n_g;
        float n_h, n_c;
        load_batch(0, n_site, n_g, n_h, n_c);
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            if ((i & 31) == 0) {
                b_site = n_site; b_g = n_g; b_h = n_h; b_c = n_c;
                load_batch(i + 32, n_site, n_g, n_h, n_c);
            }
            const int site = __shfl_sync(0xFFFFFFFFu, b_site, i & 31);
            const int g = __shfl_sync(0xFFFFFFFFu, b_g, i & 31);
            const float hv = __shfl_sync(0xFFFFFFFFu, b_h, i & 31);
            const float cg = __shfl_sync(0xFFFFFFFFu, b_c, i & 31);
            if (!INJECT && (i & 3) == 0) {
                const uint4 x = philox4x32_10(
                    make_uint4((uint32_t)rep, (uint32_t)sa, (uint32_t)(sa >> 32), (uint32_t)(i >> 2)), key);
                const uint32_t vv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float u = u01(vv[e]);
                    th4[e] = (a.rule == 0) ? -__logf(u) * Tm : 0.5f * Tm * (__logf(u) - __logf(1.0f - u));
                }
            }
            const uint32_t w = word[site];
            const bool upb = (w >> lane) & 1u;
            const int sp = upb ? 1 : -1;
            const int Sg = sum[g * 32 + lane];
            // same arithmetic as a sequential fp32 accumulation of the row: c_g * (S_g - s_i) is
            // exact for the integer-valued couplings of the penalty encodings; then + h_i
            const float f = fmaf(cg, (float)(Sg - sp), hv);
            bool flip;
            if (!INJECT) {
                const float th = th4[i & 3];
                if (a.rule == 0) {
                    const float x = upb ? 2.0f * f : -2.0f * f;  // dE = 2 s f
                    flip = x < th;
                } else {
                    flip = ((f > th) != upb);
                }
            } else {
                const float u = active ? up[i] : 1.0f;
                if (a.rule == 0) {
                    const float x = upb ? 2.0f * f : -2.0f * f;
                    flip = (x <= 0.0f) || (u < expf((float)(-(double)x / dT)));
                } else {
                    const float arg = (a.rule == 1) ? (float)(-2.0 * (double)f / dT)
                                                    : (float)(-2.0 * (1.0 / dT) * (double)f);
                    const float p_up = 1.0f / (1.0f + expf(arg));
                    flip = ((u < p_up) != upb);
                }
            }
            flip = flip && active;
            if (flip) {
                sum[g * 32 + lane] = (short)(Sg - 2 * sp);
                cur_e += upb ? 2.0f * f : -2.0f * f;
                ++n_acc;
            }
            const uint32_t fm = __ballot_sync(0xFFFFFFFFu, flip);
            if (fm && lane == 0) word[site] = w ^ fm;
            __syncwarp();
        }
        if (active && a.energy_trace) a.energy_trace[(size_t)s * a.R + rep] = cur_e;
        const bool improved = active && a.track_best && cur_e < best_e;
        if (improved) best_e = cur_e;
        const uint32_t im = __ballot_sync(0xFFFFFFFFu, improved);
        if (im) {
            uint32_t* bw = m.best_words + (size_t)blockIdx.x * n;
            for (int i = lane; i < n; i += 32) bw[i] = (bw[i] & ~im) | (word[i] & im);
        }
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) gw[i] = word[i];
    if (active) {
        a.energy[rep] = cur_e;
        if (a.track_best) a.best_energy[rep] = best_e;
        a.accepted[rep] += (unsigned long long)n_acc;
    }
}

// E_r = -1/2 sum_g c_g (S_g^2 - n_g) - sum_i h_i s_i, one warp per 32 replicas (double accumulation)
__global__ void __launch_bounds__(32)
groups_energy_kernel(const GrpDev m, const uint32_t* __restrict__ words_all, int n, int R,
                     float* __restrict__ energy) {
    extern __shared__ __align__(16) unsigned char smem[];
    int* sum = reinterpret_cast<int*>(smem);        // [n_groups][32]
    int* cnt = sum + (size_t)m.n_groups * 32;       // [n_groups]
    const int lane = threadIdx.x;
    const int rep = blockIdx.x * 32 + lane;
    const uint32_t* gw = words_all + (size_t)blockIdx.x * n;
    for (int g = 0; g < m.n_groups; ++g) sum[g * 32 + lane] = 0;
    for (int g = lane; g < m.n_groups; g += 32) cnt[g] = 0;
    __syncwarp();
    double hs = 0.0;
    for (int i = 0; i < n; ++i) {
        const int g = m.group_of[i];
        const int sp = ((gw[i] >> lane) & 1u) ? 1 : -1;
        sum[g * 32 + lane] += sp;
        if (lane == 0) cnt[g] += 1;
        hs += (double)m.h[i] * sp;
    }
    __syncwarp();
    double acc = 0.0;
    for (int g = 0; g < m.n_groups; ++g) {
        const double S = (double)sum[g * 32 + lane];
        acc += (double)m.coupling[g] * (S * S - (double)cnt[g]);
    }
    if (rep < R) energy[rep] = (float)(-0.5 * acc - hs);
}

}  // namespace

size_t groups_smem_bytes(int n, int n_groups) { return (size_t)n * 4 + (size_t)n_groups * 64; }

cudaError_t launch_sweep_groups(const GrpDev& m, const SweepDev& a, bool inject, const int* sites,
                                cudaStream_t st) {
    const size_t smem = groups_smem_bytes(a.n, m.n_groups);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e;
    const int blocks = (a.R + 31) / 32;
    if (inject) {
        e = cudaFuncSetAttribute(sweep_groups_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sweep_groups_kernel<true><<<blocks, 32, smem, st>>>(m, a, sites);
    } else {
        e = cudaFuncSetAttribute(sweep_groups_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sweep_groups_kernel<false><<<blocks, 32, smem, st>>>(m, a, sites);
    }
    return cudaGetLastError();
}

cudaError_t launch_groups_energy(const GrpDev& m, const uint32_t* words, int n, int R, float* energy,
                                 cudaStream_t st) {
    const size_t smem = (size_t)m.n_groups * 33 * 4;
    cudaError_t e = cudaFuncSetAttribute(groups_energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    groups_energy_kernel<<<(R + 31) / 32, 32, smem, st>>>(m, words, n, R, energy);
    return cudaGetLastError();
}

}  // namespace sg
