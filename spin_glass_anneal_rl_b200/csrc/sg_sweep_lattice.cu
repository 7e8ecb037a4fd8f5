// sg_sweep_lattice.cu -- K1-LAT: checkerboard multi-spin-coded sweep for 2D +-J lattices
// (BASELINE cfg2: Edwards-Anderson L = 256, 4096 replicas; instances as built by the reference's
// generator research/experimental_validation.py:134-180).
//
// Same accept rules as SpinDynamics._metropolis_update / _glauber_update / _heat_bath_update
// (reference core/spin_dynamics.py:131-191); the site order is the checkerboard sequence (all
// sites with (x + y) even in row-major order, then all odd ones), which is a legal explicit site
// order of the sequential algorithm because same-colour sites do not interact -- so a replay with
// injected uniforms against the oracle is bit-exact (tests/test_gpu_lattice.py).
//
// Multi-spin coding: bit b of word lat[w][site] is the spin of replica 32 w + b (1 = up).  For a
// site, the four "bond unsatisfied" words x_d = s ^ s_nbr ^ neg_d are added with a bit-sliced
// adder into (u0, u1, u2); per replica s f = deg - 2u, dE = 2 (deg - 2u), all integers, so the
// kernel is exact.  One thread = one (site, word): 32 attempts.  Every replica draws its own
// Philox uniform (no shared randomness between the replicas of a word) and has its own
// temperature (a ladder can live inside a word); Metropolis compares 32-bit thresholds from a
// per-replica table, the other rules / injected uniforms take the per-bit float path.
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

namespace {

constexpr uint32_t kLatStreamTag = 0xFFFFFFFEu;

// bond code of a site: bit 2d = bond d present, bit 2d+1 = bond d negative; d: 0 up (x-1),
// 1 down (x+1), 2 left (y-1), 3 right (y+1)
__device__ __forceinline__ uint32_t unsat(uint32_t s, uint32_t nb, uint32_t code, int d) {
    const uint32_t present = ((code >> (2 * d)) & 1u) ? 0xFFFFFFFFu : 0u;
    const uint32_t neg = ((code >> (2 * d + 1)) & 1u) ? 0xFFFFFFFFu : 0u;
    return (s ^ nb ^ neg) & present;
}

// rank of a site inside its colour, and number of sites of colour 0 (checkerboard sequence)
__host__ __device__ inline int color_rank(int L, int x, int y) {
    return (L & 1) ? ((x * L + y) >> 1) : (x * (L >> 1) + (y >> 1));
}
__host__ __device__ inline int color0_count(int L) { return (L * L + 1) >> 1; }

template <bool INJECT>
__global__ void __launch_bounds__(256)
lat_update_kernel(const LatDev m, const SweepDev a, int s, int color) {
    __shared__ uint32_t thr_s[32][5];      // Metropolis thresholds for s f = 1..4, per replica
    __shared__ float T_s[32];
    __shared__ unsigned int acc_s[32];
    const int L = m.L, n = L * L;
    const int w = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const int rep0 = w * 32;
    if (tid < 32) {
        const int rep = rep0 + tid;
        float T = 1.0f;
        if (rep < a.R) T = (float)a.temps[(long long)s * a.t_ss + (long long)rep * a.t_rs];
        T_s[tid] = T;
        acc_s[tid] = 0u;
        for (int k = 1; k <= 4; ++k) {
            const double p = exp(-2.0 * (double)k / (double)T);   // dE = 2 k
            thr_s[tid][k] = (p >= 1.0) ? 0xFFFFFFFFu : (uint32_t)(p * 4294967296.0);
        }
        thr_s[tid][0] = 0xFFFFFFFFu;
    }
    __syncthreads();
    // sites of this colour, in checkerboard-sequence order
    const int ncol = color == 0 ? color0_count(L) : n - color0_count(L);
    const int idx = blockIdx.x * blockDim.x + tid;
    uint32_t flipmask = 0;
    if (idx < ncol) {
        int x, y;
        if (L & 1) {
            const int k = 2 * idx + color;
            x = k / L; y = k - x * L;
        } else {
            const int half = L >> 1;
            x = idx / half;
            y = 2 * (idx - x * half) + ((x + color) & 1);
        }
        const int site = x * L + y;
        uint32_t* lat = m.lat + (size_t)w * n;
        const uint32_t sp = lat[site];
        const uint32_t code = m.bond[site];
        // neighbours with periodic addressing; a bond that does not exist (open boundary) is
        // masked out by its "present" bit
        const uint32_t xu_w = unsat(sp, lat[x > 0 ? site - L : site + n - L], code, 0);
        const uint32_t xd = unsat(sp, lat[x + 1 < L ? site + L : site + L - n], code, 1);
        const uint32_t xl = unsat(sp, lat[y > 0 ? site - 1 : site + L - 1], code, 2);
        const uint32_t xr = unsat(sp, lat[y + 1 < L ? site + 1 : site + 1 - L], code, 3);
        const uint32_t s1 = xu_w ^ xd, c1 = xu_w & xd, s2 = xl ^ xr, c2 = xl & xr;
        const uint32_t u0 = s1 ^ s2, c3 = s1 & s2;
        const uint32_t u1 = c1 ^ c2 ^ c3, u2 = (c1 & c2) | (c3 & (c1 ^ c2));
        const int deg = (int)((code & 1u) + ((code >> 2) & 1u) + ((code >> 4) & 1u) + ((code >> 6) & 1u));
        const int iseq = (color == 0 ? 0 : color0_count(L)) + color_rank(L, x, y);
        const unsigned long long sa = a.sweep_base + (unsigned long long)s;
        const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            uint32_t rr[4] = {0u, 0u, 0u, 0u};
            if (!INJECT) {
                const uint4 r4 = philox4x32_10(
                    make_uint4(kLatStreamTag - (uint32_t)q, (uint32_t)sa, (uint32_t)(sa >> 32) ^ ((uint32_t)(w + (a.rep_base >> 5)) << 8),
                               (uint32_t)(iseq)), key);
                rr[0] = r4.x; rr[1] = r4.y; rr[2] = r4.z; rr[3] = r4.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int b = q * 4 + e;
                const int rep = rep0 + b;
                const int u = (int)((u0 >> b) & 1u) + 2 * (int)((u1 >> b) & 1u) + 4 * (int)((u2 >> b) & 1u);
                const int msf = deg - 2 * u;            // s_i * f_i  (dE = 2 msf)
                const bool up = (sp >> b) & 1u;
                bool flip;
                if (!INJECT && a.rule == 0) {
                    flip = (msf <= 0) || (rr[e] < thr_s[b][msf]);
                } else {
                    const float f = up ? (float)msf : -(float)msf;   // local field
                    float u01v;
                    if (INJECT)
                        u01v = (rep < a.R) ? a.uniforms[((size_t)rep * a.n_sweeps + s) * n + iseq] : 1.0f;
                    else
                        u01v = u01(rr[e]);
                    const double dT = (double)T_s[b];
                    if (a.rule == 0) {
                        const float xx = 2.0f * (float)msf;
                        flip = (xx <= 0.0f) || (u01v < expf((float)(-(double)xx / dT)));
                    } else {
                        const float arg = (a.rule == 1) ? (float)(-2.0 * (double)f / dT)
                                                        : (float)(-2.0 * (1.0 / dT) * (double)f);
                        const float p_up = 1.0f / (1.0f + expf(arg));
                        flip = ((u01v < p_up) != up);
                    }
                }
                if (flip && rep < a.R) flipmask |= 1u << b;
            }
        }
        lat[site] = sp ^ flipmask;
    }
    // accepted flips per replica: lane b collects bit b of the warp's 32 flip masks
    unsigned int mine = 0;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
        const unsigned int c = __popc(__ballot_sync(0xFFFFFFFFu, (flipmask >> b) & 1u));
        if (lane == b) mine = c;
    }
    if (mine) atomicAdd(&acc_s[lane], mine);
    __syncthreads();
    if (tid < 32 && rep0 + tid < a.R && acc_s[tid])
        atomicAdd(&a.accepted[rep0 + tid], (unsigned long long)acc_s[tid]);
}

// per-replica energy E = 2 * (#unsatisfied bonds) - (#bonds): one block per word, vertical
// (bit-sliced) counters per thread, then per-bit extraction and a shared-memory reduction
__global__ void __launch_bounds__(256)
lat_energy_kernel(const LatDev m, const uint32_t* __restrict__ lat_all, int R, int n_bonds,
                  float* __restrict__ energy) {
    __shared__ int tot[32];
    const int L = m.L, n = L * L, w = blockIdx.x, tid = threadIdx.x;
    const uint32_t* lat = lat_all + (size_t)w * n;
    if (tid < 32) tot[tid] = 0;
    __syncthreads();
    uint32_t c[18];
#pragma unroll
    for (int p = 0; p < 18; ++p) c[p] = 0u;
    auto add = [&](uint32_t xw) {
#pragma unroll
        for (int p = 0; p < 18; ++p) {
            const uint32_t t = c[p] & xw;
            c[p] ^= xw;
            xw = t;
        }
    };
    for (int site = tid; site < n; site += blockDim.x) {
        const int x = site / L, y = site - x * L;
        const uint32_t sp = lat[site], code = m.bond[site];
        // forward bonds only: down and right
        add(unsat(sp, x + 1 < L ? lat[site + L] : lat[site + L - n], code, 1));
        add(unsat(sp, y + 1 < L ? lat[site + 1] : lat[site + 1 - L], code, 3));
    }
    for (int b = 0; b < 32; ++b) {
        int v = 0;
#pragma unroll
        for (int p = 0; p < 18; ++p) v += (int)((c[p] >> b) & 1u) << p;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((tid & 31) == 0 && v) atomicAdd(&tot[b], v);
    }
    __syncthreads();
    if (tid < 32 && w * 32 + tid < R) energy[w * 32 + tid] = (float)(2 * tot[tid] - n_bonds);
}

// best tracking: replicas whose energy improved copy their bit into the best lattice
__global__ void lat_best_kernel(const uint32_t* __restrict__ lat, uint32_t* __restrict__ best_lat, int n,
                                int R, const float* __restrict__ energy,
                                const float* __restrict__ best_energy) {
    __shared__ uint32_t mask_s;
    const int w = blockIdx.y;
    if (threadIdx.x < 32) {
        const int rep = w * 32 + threadIdx.x;
        const bool imp = rep < R && energy[rep] < best_energy[rep];
        const uint32_t mk = __ballot_sync(0xFFFFFFFFu, imp);
        if (threadIdx.x == 0) mask_s = mk;
    }
    __syncthreads();
    const uint32_t mk = mask_s;
    if (mk) {
        for (int site = blockIdx.x * blockDim.x + threadIdx.x; site < n; site += gridDim.x * blockDim.x) {
            const size_t o = (size_t)w * n + site;
            best_lat[o] = (best_lat[o] & ~mk) | (lat[o] & mk);
        }
    }
}

// energies are final for this sweep: update best energies (after every block of lat_best_kernel
// has read them) and write the trace row
__global__ void lat_best_energy_kernel(int R, const float* __restrict__ energy,
                                       float* __restrict__ best_energy, float* __restrict__ trace_row,
                                       int track_best) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float e = energy[r];
    if (track_best && e < best_energy[r]) best_energy[r] = e;
    if (trace_row) trace_row[r] = e;
}

// spins [R][n] int8 <-> lat[W][n] bit planes
__global__ void lat_pack_kernel(const int8_t* __restrict__ spins, int n, int R, uint32_t* __restrict__ lat) {
    const int w = blockIdx.y;
    for (int site = blockIdx.x * blockDim.x + threadIdx.x; site < n; site += gridDim.x * blockDim.x) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int rep = w * 32 + b;
            const bool up = rep < R ? spins[(size_t)rep * n + site] > 0 : true;
            bits |= (up ? 1u : 0u) << b;
        }
        lat[(size_t)w * n + site] = bits;
    }
}

__global__ void lat_unpack_kernel(const uint32_t* __restrict__ lat, int n, int R, int8_t* __restrict__ spins) {
    const int w = blockIdx.y;
    for (int site = blockIdx.x * blockDim.x + threadIdx.x; site < n; site += gridDim.x * blockDim.x) {
        const uint32_t bits = lat[(size_t)w * n + site];
        for (int b = 0; b < 32; ++b) {
            const int rep = w * 32 + b;
            if (rep < R) spins[(size_t)rep * n + site] = ((bits >> b) & 1u) ? 1 : -1;
        }
    }
}

}  // namespace

int lattice_sequence_index(int L, int x, int y) {
    const int color = (x + y) & 1;
    return (color == 0 ? 0 : color0_count(L)) + color_rank(L, x, y);
}

cudaError_t launch_lat_pack(const int8_t* spins, int n, int R, uint32_t* lat, cudaStream_t st) {
    dim3 grid((n + 255) / 256 < 64 ? (n + 255) / 256 : 64, (R + 31) / 32);
    lat_pack_kernel<<<grid, 256, 0, st>>>(spins, n, R, lat);
    return cudaGetLastError();
}
cudaError_t launch_lat_unpack(const uint32_t* lat, int n, int R, int8_t* spins, cudaStream_t st) {
    dim3 grid((n + 255) / 256 < 64 ? (n + 255) / 256 : 64, (R + 31) / 32);
    lat_unpack_kernel<<<grid, 256, 0, st>>>(lat, n, R, spins);
    return cudaGetLastError();
}
cudaError_t launch_lat_energy(const LatDev& m, const uint32_t* lat, int R, float* energy, cudaStream_t st) {
    lat_energy_kernel<<<(R + 31) / 32, 256, 0, st>>>(m, lat, R, m.n_bonds, energy);
    return cudaGetLastError();
}

// n_sweeps checkerboard sweeps; per sweep: colour 0, colour 1, (energy, best) when asked
cudaError_t launch_sweep_lattice(const LatDev& m, const SweepDev& a, bool inject, uint64_t* launches,
                                 cudaStream_t st) {
    const int L = m.L, n = L * L, W = (a.R + 31) / 32;
    for (int s = 0; s < a.n_sweeps; ++s) {
        for (int color = 0; color < 2; ++color) {
            const int ncol = color == 0 ? color0_count(L) : n - color0_count(L);
            if (ncol == 0) continue;
            dim3 grid((ncol + 255) / 256, W);
            if (inject) lat_update_kernel<true><<<grid, 256, 0, st>>>(m, a, s, color);
            else lat_update_kernel<false><<<grid, 256, 0, st>>>(m, a, s, color);
            ++*launches;
        }
        const bool last = (s + 1 == a.n_sweeps);
        if (a.track_best || a.energy_trace || last) {
            lat_energy_kernel<<<W, 256, 0, st>>>(m, m.lat, a.R, m.n_bonds, a.energy);
            ++*launches;
            if (a.track_best) {
                dim3 grid(32, W);
                lat_best_kernel<<<grid, 256, 0, st>>>(m.lat, m.best_lat, n, a.R, a.energy, a.best_energy);
                ++*launches;
            }
            lat_best_energy_kernel<<<(a.R + 255) / 256, 256, 0, st>>>(
                a.R, a.energy, a.best_energy,
                a.energy_trace ? a.energy_trace + (size_t)s * a.R : nullptr, a.track_best);
            ++*launches;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

}  // namespace sg
