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
// Two kernels.  The default (second half of this file) deals the groups out to P partitions and
// runs a P x W grid of one-warp CTAs, one launch per sweep.  The first one, below, is the simple
// form it grew out of (kept for single-group models and as a cross-check, SG_GRP_PART=0):
// one warp = 32 replicas (lane = replica) with its whole state in shared memory: bit b of
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
                    make_uint4((uint32_t)(rep + a.rep_base), (uint32_t)sa, (uint32_t)(sa >> 32), (uint32_t)(i >> 2)), key);
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

// ================================================================ partitioned variant (K1-GRP/P)
//
// Groups do not interact, so attempts on different groups commute exactly: the state after a sweep
// does not depend on how attempts of different groups interleave, only on the order inside each
// group (and on using the random numbers of the same attempt index).  With few replica words
// (cfg5: 1024 replicas = 32 warps, one per SM, each latency-bound on its own dependency chain)
// the groups are therefore dealt out to P partitions and the grid becomes P x W one-warp CTAs,
// each holding only its partition's state in shared memory (several resident per SM).  Per
// sweep: a prepass orders the attempt list by partition (stable), one launch runs every (partition,
// replica word) pair, a finalize step adds the partitions' energy changes per replica, keeps the
// per-sweep trace and the best-so-far configurations.

// stable counting sort of one sweep's attempts by the partition of their site:
// out[s][aoff[s][p] + k] = {attempt index, site} of the k-th attempt of partition p
__global__ void __launch_bounds__(1024)
grp_order_attempts_kernel(const int* __restrict__ sites_g, int n, const int* __restrict__ group_of,
                          const int* __restrict__ part_of_group, int P, int2* __restrict__ out,
                          int* __restrict__ aoff) {
    __shared__ int count[32];          // attempts per partition (P <= 32)
    __shared__ int base[32];           // running output offset per partition
    __shared__ int wtot[32][32];       // [warp][partition] counts of the current chunk
    const int s = blockIdx.x;
    const int* tab = sites_g + (size_t)s * n;
    int2* o = out + (size_t)s * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 32) count[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) atomicAdd(&count[part_of_group[group_of[tab[i]]]], 1);
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int p = 0; p < P; ++p) {
            base[p] = acc;
            aoff[(size_t)s * (P + 1) + p] = acc;
            acc += count[p];
        }
        aoff[(size_t)s * (P + 1) + P] = acc;
    }
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + tid;
        const bool valid = i < n;
        const int site = valid ? tab[i] : 0;
        const int part = valid ? part_of_group[group_of[site]] : -1;
        wtot[warp][lane] = 0;
        __syncwarp();
        const unsigned same = __match_any_sync(0xFFFFFFFFu, part);
        const int rank = __popc(same & ((1u << lane) - 1u));
        if (valid && rank == 0) wtot[warp][part] = __popc(same);
        __syncthreads();
        if (valid) {
            int before = 0;
            for (int w2 = 0; w2 < warp; ++w2) before += wtot[w2][part];
            o[base[part] + before + rank] = make_int2(i, site);
        }
        __syncthreads();
        if (tid < P) {
            int t = 0;
            for (int w2 = 0; w2 < 32; ++w2) t += wtot[w2][tid];
            base[tid] += t;
        }
        __syncthreads();
    }
}

// group sums of every replica word: sums[w][g][lane] (int16), one warp per (group, word)
__global__ void __launch_bounds__(32)
grp_sums_kernel(const uint32_t* __restrict__ words, int n, int n_groups,
                const int* __restrict__ goff, const int* __restrict__ gsites, short* __restrict__ sums) {
    const int g = blockIdx.x, w = blockIdx.y, lane = threadIdx.x;
    const uint32_t* gw = words + (size_t)w * n;
    int acc = 0;
    for (int k = goff[g]; k < goff[g + 1]; ++k) acc += ((gw[gsites[k]] >> lane) & 1u) ? 1 : -1;
    sums[((size_t)w * n_groups + g) * 32 + lane] = (short)acc;
}

template <bool INJECT>
__global__ void __launch_bounds__(32)
sweep_groups_part_kernel(const GrpDev m, const GrpPartDev q, const SweepDev a, const int s,
                         const int2* __restrict__ alist, const int* __restrict__ aoff,
                         double* __restrict__ de_part) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int p = blockIdx.x, w = blockIdx.y, lane = threadIdx.x;
    const int n = a.n;
    const int s_off = q.part_off[p], ns = q.part_off[p + 1] - s_off;
    const int g_off = q.part_goff[p], ng = q.part_goff[p + 1] - g_off;
    uint32_t* word = reinterpret_cast<uint32_t*>(smem);                       // [ns]
    short* sum = reinterpret_cast<short*>(smem + (size_t)q.max_sites * 4);    // [ng][32]
    const int rep = w * 32 + lane;
    const bool active = rep < a.R;
    uint32_t* gw = m.words + (size_t)w * n;
    short* gs = q.sums + ((size_t)w * m.n_groups) * 32;
    for (int k = lane; k < ns; k += 32) word[k] = gw[q.part_sites[s_off + k]];
    for (int k = 0; k < ng; ++k) sum[k * 32 + lane] = gs[(size_t)q.part_groups[g_off + k] * 32 + lane];
    __syncwarp();

    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const unsigned long long sa = a.sweep_base + (unsigned long long)s;
    const double dT = active ? a.temps[(long long)s * a.t_ss + (long long)rep * a.t_rs] : 1.0;
    const float Tm = (float)dT;
    const float* up = INJECT ? a.uniforms + ((size_t)rep * a.n_sweeps + s) * n : nullptr;
    const int2* al = alist + (size_t)s * n + aoff[(size_t)s * (q.P + 1) + p];
    const int na = aoff[(size_t)s * (q.P + 1) + p + 1] - aoff[(size_t)s * (q.P + 1) + p];
    double de = 0.0;
    unsigned int n_acc = 0;

    // per-attempt data that does not depend on the state, one attempt per lane, a batch ahead
    int b_i = 0, b_ls = 0, b_lg = 0;
    float b_h = 0.0f, b_c = 0.0f;
    auto load_batch = [&](int j0, int& li, int& ls, int& lg, float& lh, float& lc) {
        const int j = j0 + lane;
        li = 0; ls = 0; lg = 0; lh = 0.0f; lc = 0.0f;
        if (j < na) {
            const int2 e = al[j];
            const int g = m.group_of[e.y];
            li = e.x;
            ls = q.local_site[e.y];
            lg = q.local_group[g];
            lh = m.h[e.y];
            lc = m.coupling[g];
        }
    };
    auto threshold = [&](int i) -> float {
        const uint4 x = philox4x32_10(
            make_uint4((uint32_t)(rep + a.rep_base), (uint32_t)sa, (uint32_t)(sa >> 32), (uint32_t)(i >> 2)), key);
        const uint32_t sel = (i & 2) ? ((i & 1) ? x.w : x.z) : ((i & 1) ? x.y : x.x);
        const float u = u01(sel);
        return (a.rule == 0) ? -__logf(u) * Tm : 0.5f * Tm * (__logf(u) - __logf(1.0f - u));
    };
    int n_i, n_ls, n_lg;
    float n_h, n_c;
    load_batch(0, n_i, n_ls, n_lg, n_h, n_c);
    float th_next = 0.0f;
    if (!INJECT && na > 0) th_next = threshold(__shfl_sync(0xFFFFFFFFu, n_i, 0));
#pragma unroll 1
    for (int j = 0; j < na; ++j) {
        if ((j & 31) == 0) {
            b_i = n_i; b_ls = n_ls; b_lg = n_lg; b_h = n_h; b_c = n_c;
            load_batch(j + 32, n_i, n_ls, n_lg, n_h, n_c);
        }
        const int i = __shfl_sync(0xFFFFFFFFu, b_i, j & 31);
        const int ls = __shfl_sync(0xFFFFFFFFu, b_ls, j & 31);
        const int lg = __shfl_sync(0xFFFFFFFFu, b_lg, j & 31);
        const float hv = __shfl_sync(0xFFFFFFFFu, b_h, j & 31);
        const float cg = __shfl_sync(0xFFFFFFFFu, b_c, j & 31);
        const float th = th_next;
        if (!INJECT && j + 1 < na) {
            // the next attempt's threshold is independent of the state: its Philox rounds overlap
            // this attempt's shared-memory round trips
            const int i_next = ((j + 1) & 31) ? __shfl_sync(0xFFFFFFFFu, b_i, (j + 1) & 31)
                                              : __shfl_sync(0xFFFFFFFFu, n_i, 0);
            th_next = threshold(i_next);
        }
        const uint32_t wv = word[ls];
        const bool upb = (wv >> lane) & 1u;
        const int sp = upb ? 1 : -1;
        const int Sg = sum[lg * 32 + lane];
        const float f = fmaf(cg, (float)(Sg - sp), hv);
        bool flip;
        if (!INJECT) {
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
            sum[lg * 32 + lane] = (short)(Sg - 2 * sp);
            de += (double)(upb ? 2.0f * f : -2.0f * f);
            ++n_acc;
        }
        const uint32_t fm = __ballot_sync(0xFFFFFFFFu, flip);
        if (fm && lane == 0) word[ls] = wv ^ fm;
        __syncwarp();
    }
    __syncwarp();
    for (int k = lane; k < ns; k += 32) gw[q.part_sites[s_off + k]] = word[k];
    for (int k = 0; k < ng; ++k) gs[(size_t)q.part_groups[g_off + k] * 32 + lane] = sum[k * 32 + lane];
    if (active) {
        de_part[(size_t)p * a.R + rep] = de;
        if (n_acc) atomicAdd(&a.accepted[rep], (unsigned long long)n_acc);
    }
}

// after a sweep: E += sum of the partitions' changes (fixed order), trace, best-so-far mask
__global__ void __launch_bounds__(32)
grp_finalize_kernel(const SweepDev a, int s, int P, const double* __restrict__ de_part,
                    uint32_t* __restrict__ improved_mask) {
    const int w = blockIdx.x, lane = threadIdx.x;
    const int rep = w * 32 + lane;
    bool improved = false;
    if (rep < a.R) {
        double acc = 0.0;
        for (int p = 0; p < P; ++p) acc += de_part[(size_t)p * a.R + rep];
        const float e = (float)((double)a.energy[rep] + acc);
        a.energy[rep] = e;
        if (a.energy_trace) a.energy_trace[(size_t)s * a.R + rep] = e;
        if (a.track_best && e < a.best_energy[rep]) {
            a.best_energy[rep] = e;
            improved = true;
        }
    }
    const uint32_t im = __ballot_sync(0xFFFFFFFFu, improved);
    if (lane == 0) improved_mask[w] = im;
}

__global__ void __launch_bounds__(256)
grp_keep_best_kernel(const uint32_t* __restrict__ words, uint32_t* __restrict__ best_words, int n,
                     const uint32_t* __restrict__ improved_mask) {
    const int w = blockIdx.y;
    const uint32_t im = improved_mask[w];
    if (im == 0u) return;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const size_t o = (size_t)w * n + i;
    best_words[o] = (best_words[o] & ~im) | (words[o] & im);
}

}  // namespace

size_t groups_smem_bytes(int n, int n_groups) { return (size_t)n * 4 + (size_t)n_groups * 64; }

size_t groups_part_scratch_bytes(int n, int n_sweeps, int R, int P) {
    const size_t W = ((size_t)R + 31) / 32;
    size_t b = (size_t)n_sweeps * n * sizeof(int2);                  // ordered attempts
    b += ((size_t)n_sweeps * (P + 1) * sizeof(int) + 15) & ~(size_t)15;   // offsets
    b += (size_t)P * R * sizeof(double);                              // energy changes
    b += (W * sizeof(uint32_t) + 15) & ~(size_t)15;                   // improved masks
    return b;
}

cudaError_t launch_sweep_groups_part(const GrpDev& m, const GrpPartDev& q, const SweepDev& a, bool inject,
                                     const int* sites, void* scratch, uint64_t* launches,
                                     cudaStream_t st) {
    const int n = a.n, P = q.P;
    const int W = (a.R + 31) / 32;
    unsigned char* sp = static_cast<unsigned char*>(scratch);
    int2* alist = reinterpret_cast<int2*>(sp);
    sp += (size_t)a.n_sweeps * n * sizeof(int2);
    int* aoff = reinterpret_cast<int*>(sp);
    sp += ((size_t)a.n_sweeps * (P + 1) * sizeof(int) + 15) & ~(size_t)15;
    double* de_part = reinterpret_cast<double*>(sp);
    sp += (size_t)P * a.R * sizeof(double);
    uint32_t* imask = reinterpret_cast<uint32_t*>(sp);

    grp_order_attempts_kernel<<<a.n_sweeps, 1024, 0, st>>>(sites, n, m.group_of, q.part_of_group, P, alist, aoff);
    grp_sums_kernel<<<dim3((unsigned)m.n_groups, (unsigned)W), 32, 0, st>>>(m.words, n, m.n_groups, q.goff,
                                                                            q.gsites, q.sums);
    *launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)q.max_sites * 4 + (size_t)q.max_groups * 64;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    if (inject)
        e = cudaFuncSetAttribute(sweep_groups_part_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else
        e = cudaFuncSetAttribute(sweep_groups_part_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const dim3 grid((unsigned)P, (unsigned)W);
    for (int s = 0; s < a.n_sweeps; ++s) {
        if (inject)
            sweep_groups_part_kernel<true><<<grid, 32, smem, st>>>(m, q, a, s, alist, aoff, de_part);
        else
            sweep_groups_part_kernel<false><<<grid, 32, smem, st>>>(m, q, a, s, alist, aoff, de_part);
        grp_finalize_kernel<<<W, 32, 0, st>>>(a, s, P, de_part, imask);
        if (a.track_best)
            grp_keep_best_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)W), 256, 0, st>>>(
                m.words, m.best_words, n, imask);
        *launches += a.track_best ? 3 : 2;
    }
    return cudaGetLastError();
}

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
