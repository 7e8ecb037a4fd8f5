// sg_sweep_csr.cu -- K1-CSR: the replica-batched Monte Carlo sweep for SPARSE couplings.
//
// Same contract as sg_sweep.cu (SpinDynamics.sweep() for R replicas, reference
// core/spin_dynamics.py:61-94,131-191) for the models the reference's callers build as sparse COO
// (problems/base.py:107-116): the 50 000-spin scheduling QUBO (block cliques, 99 couplings per
// row) and the 2D +-J lattice (4 couplings per row), which cannot be held as dense matrices.
//
// Layout: replica-minor.  spins[site][r] (int8) and fields[site][r] (fp32) with r padded to 32,
// so that the 32 replicas of a warp (lane = replica) touch one 32 / 128-byte segment per site.
// All replicas of a launch visit the sites in one order (site table built once per launch), a
// warp walks the attempts of its 32 replicas sequentially:
//   f = fields[site][lane]; accept test (Philox thresholds or injected uniforms);
//   for every coupling (j, v) of row `site` of J^T:  fields[j][lane] += delta * v   (delta = -2 s or 0)
// The loads of the next attempt are issued before the current row is applied and patched if the
// next site is one of its neighbours.  The energy is carried incrementally,
// E += 2 s f - 2 J_ss (exact for integer couplings), and compared with the best after every sweep.
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

namespace {

constexpr int kCsrWarps = 4;  // warps per block, each owns 32 replicas

__global__ void csr_sites_kernel(int mode, unsigned long long seed, unsigned long long sweep_base,
                                 int n, int n_sweeps, const int* __restrict__ explicit_sites,
                                 long long s_ss, int* __restrict__ out) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const int quads = (n + 3) / 4;
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
            if (i < n) {
                int site;
                if (mode == 0) site = i;
                else if (mode == 1) site = (int)v[e];
                else site = explicit_sites[(long long)s * s_ss + i];
                out[(size_t)s * n + i] = site;
            }
        }
    }
}

template <bool INJECT>
__global__ void __launch_bounds__(kCsrWarps * 32)
sweep_csr_kernel(const CsrDev m, const SweepDev a, const int* __restrict__ sites_g) {
    const int lane = threadIdx.x & 31;
    const int wg = blockIdx.x * kCsrWarps + (threadIdx.x >> 5);   // replica group of this warp
    const int rep = wg * 32 + lane;
    if (wg * 32 >= a.R) return;
    const bool active = rep < a.R;
    const int n = a.n, Rp = m.Rp, n_sweeps = a.n_sweeps;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    int8_t* S = m.spins + rep;
    float* F = m.fields + rep;

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
        // prefetched state of the next attempt
        int site = tab[0];
        float f = F[(size_t)site * Rp];
        int sp = S[(size_t)site * Rp];
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const int nsite = (i + 1 < n) ? tab[i + 1] : -1;
            float fn = 0.0f;
            int spn = 1;
            if (nsite >= 0) {
                fn = F[(size_t)nsite * Rp];
                spn = S[(size_t)nsite * Rp];
            }
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
            const bool upb = sp > 0;
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
            const float delta = flip ? (upb ? -2.0f : 2.0f) : 0.0f;
            if (__any_sync(0xFFFFFFFFu, flip)) {
                const float dg = m.diag[site];
                if (flip) {
                    S[(size_t)site * Rp] = (int8_t)(-sp);
                    cur_e += (upb ? 2.0f * f : -2.0f * f) - 2.0f * dg;
                    ++n_acc;
                }
                if (nsite == site) {
                    spn = flip ? -sp : sp;
                    // own diagonal enters the field of the same site
                }
                const long long e0 = m.rowptr[site], e1 = m.rowptr[site + 1];
#pragma unroll 4
                for (long long e = e0; e < e1; ++e) {
                    const int j = m.colidx[e];
                    const float v = m.val[e];
                    float* pf = F + (size_t)j * Rp;
                    *pf = fmaf(delta, v, *pf);
                    if (j == nsite) fn = fmaf(delta, v, fn);
                }
            }
            site = nsite;
            f = fn;
            sp = spn;
        }
        if (!m.symmetric) {
            // asymmetric couplings: dE of a flip also involves the column of J, so recompute
            // E = -1/2 sum_j s_j (f_j + h_j) from the resident fields, as the dense kernels do
            float acc = 0.0f;
            for (int j = 0; j < n; ++j) {
                const float t = F[(size_t)j * Rp] + m.h[j];
                acc += (S[(size_t)j * Rp] > 0) ? t : -t;
            }
            cur_e = -0.5f * acc;
        }
        if (active) {
            if (a.energy_trace) a.energy_trace[(size_t)s * a.R + rep] = cur_e;
        }
        const bool improved = active && a.track_best && cur_e < best_e;
        if (improved) best_e = cur_e;
        if (__any_sync(0xFFFFFFFFu, improved)) {
            int8_t* B = m.best_spins + rep;
            for (int j = 0; j < n; ++j)
                if (improved) B[(size_t)j * Rp] = S[(size_t)j * Rp];
        }
    }
    if (active) {
        a.energy[rep] = cur_e;
        if (a.track_best) a.best_energy[rep] = best_e;
        a.accepted[rep] += (unsigned long long)n_acc;
    }
}

// F[j][r] = h_j + sum_e val[e] * S[col[e]][r] over row j of J (one thread per (j, r))
__global__ void csr_fields_kernel(const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                                  const float* __restrict__ val, const float* __restrict__ h, int n,
                                  int Rp, const int8_t* __restrict__ S, float* __restrict__ F) {
    const size_t total = (size_t)n * Rp;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx / Rp), r = (int)(idx - (size_t)j * Rp);
        float acc = 0.0f;
        for (long long e = rowptr[j]; e < rowptr[j + 1]; ++e)
            acc = fmaf(val[e], (float)S[(size_t)colidx[e] * Rp + r], acc);
        F[idx] = acc + h[j];
    }
}

// E_r = -1/2 sum_j S[j][r] (F[j][r] + h_j)   (one thread per replica, coalesced over r)
__global__ void csr_energy_kernel(const float* __restrict__ h, int n, int Rp, int R,
                                  const int8_t* __restrict__ S, const float* __restrict__ F,
                                  float* __restrict__ energy) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    double acc = 0.0;
    for (int j = 0; j < n; ++j) {
        const float t = F[(size_t)j * Rp + r] + h[j];
        acc += (S[(size_t)j * Rp + r] > 0) ? (double)t : -(double)t;
    }
    energy[r] = (float)(-0.5 * acc);
}

// [R][n] row-major <-> [n][Rp] replica-minor
template <typename T>
__global__ void to_replica_minor_kernel(const T* __restrict__ src, int n, int R, int Rp,
                                        T* __restrict__ dst, T pad) {
    __shared__ T tile[32][33];
    const int j0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int r = r0 + k, j = j0 + tx;
        tile[k][tx] = (r < R && j < n) ? src[(size_t)r * n + j] : pad;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int j = j0 + k, r = r0 + tx;
        if (j < n && r < Rp) dst[(size_t)j * Rp + r] = tile[tx][k];
    }
}

template <typename T>
__global__ void from_replica_minor_kernel(const T* __restrict__ src, int n, int R, int Rp,
                                          T* __restrict__ dst) {
    __shared__ T tile[32][33];
    const int j0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int k = ty; k < 32; k += 8) {
        const int j = j0 + k, r = r0 + tx;
        tile[k][tx] = (j < n && r < Rp) ? src[(size_t)j * Rp + r] : T(0);
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int r = r0 + k, j = j0 + tx;
        if (r < R && j < n) dst[(size_t)r * n + j] = tile[tx][k];
    }
}

}  // namespace

size_t csr_sites_bytes(int n, int n_sweeps) { return (size_t)n_sweeps * n * sizeof(int); }

cudaError_t launch_sites_table(const SweepDev& a, int* out, cudaStream_t st) {
    const long long total = (long long)a.n_sweeps * ((a.n + 3) / 4);
    int grid = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
    csr_sites_kernel<<<grid, 256, 0, st>>>(a.site_mode, a.seed, a.sweep_base, a.n, a.n_sweeps,
                                           a.sites, a.s_ss, out);
    return cudaGetLastError();
}

cudaError_t launch_sweep_csr(const CsrDev& m, const SweepDev& a, bool inject, void* sites_buf,
                             cudaStream_t st) {
    int* sites = static_cast<int*>(sites_buf);
    cudaError_t e = launch_sites_table(a, sites, st);
    if (e != cudaSuccess) return e;
    const int groups = (a.R + 31) / 32;
    const int blocks = (groups + kCsrWarps - 1) / kCsrWarps;
    if (inject)
        sweep_csr_kernel<true><<<blocks, kCsrWarps * 32, 0, st>>>(m, a, sites);
    else
        sweep_csr_kernel<false><<<blocks, kCsrWarps * 32, 0, st>>>(m, a, sites);
    return cudaGetLastError();
}

cudaError_t launch_csr_fields(const CsrDev& m, const long long* rowptr_rows, const int* colidx_rows,
                              const float* val_rows, const float* h, int n, int R, float* energy,
                              cudaStream_t st) {
    const size_t total = (size_t)n * m.Rp;
    int grid = (int)((total + 255) / 256 < (size_t)148 * 32 ? (total + 255) / 256 : (size_t)148 * 32);
    csr_fields_kernel<<<grid, 256, 0, st>>>(rowptr_rows, colidx_rows, val_rows, h, n, m.Rp, m.spins,
                                            m.fields);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    csr_energy_kernel<<<(R + 127) / 128, 128, 0, st>>>(h, n, m.Rp, R, m.spins, m.fields, energy);
    return cudaGetLastError();
}

cudaError_t launch_to_replica_minor_i8(const int8_t* src, int n, int R, int Rp, int8_t* dst,
                                       cudaStream_t st) {
    dim3 grid((n + 31) / 32, (Rp + 31) / 32), block(32, 8);
    to_replica_minor_kernel<int8_t><<<grid, block, 0, st>>>(src, n, R, Rp, dst, (int8_t)1);
    return cudaGetLastError();
}
cudaError_t launch_from_replica_minor_i8(const int8_t* src, int n, int R, int Rp, int8_t* dst,
                                         cudaStream_t st) {
    dim3 grid((n + 31) / 32, (Rp + 31) / 32), block(32, 8);
    from_replica_minor_kernel<int8_t><<<grid, block, 0, st>>>(src, n, R, Rp, dst);
    return cudaGetLastError();
}
cudaError_t launch_from_replica_minor_f32(const float* src, int n, int R, int Rp, float* dst,
                                          cudaStream_t st) {
    dim3 grid((n + 31) / 32, (Rp + 31) / 32), block(32, 8);
    from_replica_minor_kernel<float><<<grid, block, 0, st>>>(src, n, R, Rp, dst);
    return cudaGetLastError();
}

}  // namespace sg
