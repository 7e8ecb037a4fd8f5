// sg_exchange.cu -- K3: replica exchange between adjacent rungs of a temperature ladder.
//
// Restates ParallelTempering._nearest_neighbor_exchange / _attempt_single_exchange
// (reference annealing/parallel_tempering.py:214-258):
//   for i in range(start, K-1, 2):  p = min(1, exp((b_{i+1}-b_i) (E_{i+1}-E_i)));
//   attempts[i] += 1;  if rand() < p: swap, accepts[i] += 1
// Rung 0 is the hottest.  The reference swaps the spin configurations of slots i and
// i+1; here the configurations stay where they are and the TEMPERATURES move: the
// map rung -> replica (rep_at) and the per-replica temperature are swapped instead,
// which is the same Markov chain with O(1) traffic per accepted exchange.
// One warp per ladder, partner energies by warp shuffle; pairs of one parity are disjoint.
// Sharded ladders (one ladder over several GPUs): `energy` is the all-gathered table indexed by
// GLOBAL replica id, every rank takes the same decisions from the same counter RNG and keeps the
// same rung -> replica map; only the temperatures of its own replicas are stored locally.
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

namespace {

// 53-bit uniform in [0, 1) from two Philox words
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// The temperature of a replica is stored only by the rank that holds it (sharded ladders: the
// rung -> replica map is global and kept identically by every rank, rep_temp is local).
__device__ __forceinline__ void set_temp(const ExchangeDev& a, int rep, double T) {
    const int loc = rep - a.rep_lo;
    if (loc >= 0 && loc < a.rep_n) a.rep_temp[loc] = T;
}

// One WARP per ladder: lane j owns rungs j, j + 32, ... and keeps (replica, energy) of its rung
// in registers; the upper partner of a pair comes from the next lane with a shuffle, so every
// energy is read once.  Pairs of one parity are disjoint, so all lanes decide at once.
__global__ void exchange_kernel(const ExchangeDev a) {
    const int lane = threadIdx.x & 31;
    const int l = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (l >= a.L) return;
    const int npairs = (a.K - a.parity) / 2;  // pairs (k, k+1), k = parity, parity+2, ... < K-1
    int* slot = a.rep_at + (size_t)l * a.K;
    for (int k0 = 0; k0 < a.K; k0 += 32) {
        const int k = k0 + lane;
        const int rep = (k < a.K) ? slot[k] : 0;
        const double E = (k < a.K) ? (double)a.energy[rep] : 0.0;
        // partner on rung k + 1: the next lane's values; the last lane reaches into the next chunk
        int rep_up = __shfl_down_sync(0xFFFFFFFFu, rep, 1);
        double E_up = __shfl_down_sync(0xFFFFFFFFu, E, 1);
        if (lane == 31 && k + 1 < a.K) {
            rep_up = slot[k + 1];
            E_up = (double)a.energy[rep_up];
        }
        const bool mine = k + 1 < a.K && k >= a.parity && ((k - a.parity) & 1) == 0;
        // (all loads and shuffles of the chunk are done before any lane rewrites the map)
        __syncwarp();
        if (mine) {
            const int m = (k - a.parity) >> 1;
            const double Ti = a.ladder[k], Tj = a.ladder[k + 1];
            const double arg = (1.0 / Tj - 1.0 / Ti) * (E_up - E);
            const double prob = fmin(1.0, exp(arg));
            double u;
            if (a.inject) {
                u = a.uniforms[(size_t)l * (a.K / 2) + m];
            } else {
                const int t = l * npairs + m;
                const uint4 x = philox4x32_10(
                    make_uint4(0xE8C4A46Eu, (uint32_t)a.round, (uint32_t)(a.round >> 32), (uint32_t)t),
                    make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                u = u53(x.x, x.y);
            }
            const size_t st = (size_t)l * (a.K - 1) + k;
            a.attempts[st] += 1u;
            if (u < prob) {
                slot[k] = rep_up;
                slot[k + 1] = rep;
                set_temp(a, rep, Tj);
                set_temp(a, rep_up, Ti);
                a.accepts[st] += 1u;
            }
        }
        __syncwarp();
    }
}

// ParallelTempering._all_pairs_exchange, CPU branch (reference annealing/parallel_tempering.py:
// 228-232): every pair i < j is attempted with probability 0.1, in order, each attempt seeing the
// swaps before it; statistics are filed under min(i, j).  A dependency chain over K (K-1) / 2
// scalars per ladder: one thread per ladder.  Draw c of ladder l is uniforms[l * K (K-1) + c]
// (injected: the caller's stream in consumption order) or Philox word pair c.
__global__ void exchange_all_pairs_kernel(const ExchangeDev a) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= a.L) return;
    int* slot = a.rep_at + (size_t)l * a.K;
    const double* inj = a.inject ? a.uniforms + (size_t)l * a.K * (a.K - 1) : nullptr;
    unsigned int c = 0;
    auto draw = [&]() -> double {
        double u;
        if (inj) {
            u = inj[c];
        } else {
            const uint4 x = philox4x32_10(
                make_uint4(0xA11FA125u ^ (uint32_t)l, (uint32_t)a.round, (uint32_t)(a.round >> 32), c),
                make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
            u = u53(x.x, x.y);
        }
        ++c;
        return u;
    };
    for (int i = 0; i + 1 < a.K; ++i)
        for (int j = i + 1; j < a.K; ++j) {
            if (!(draw() < 0.1)) continue;
            const int ra = slot[i], rb = slot[j];
            const double Ti = a.ladder[i], Tj = a.ladder[j];
            const double arg = (1.0 / Tj - 1.0 / Ti) * ((double)a.energy[rb] - (double)a.energy[ra]);
            const double prob = fmin(1.0, exp(arg));
            const size_t st = (size_t)l * (a.K - 1) + i;
            a.attempts[st] += 1u;
            if (draw() < prob) {
                slot[i] = rb;
                slot[j] = ra;
                set_temp(a, ra, Tj);
                set_temp(a, rb, Ti);
                a.accepts[st] += 1u;
            }
        }
}

// hit[0] = first `round` at which min_r energy[r] <= target (-1 until then), hit[1] = that replica
// (global id).  hit may live in pinned host memory: the host polls it without synchronising.
__global__ void check_target_kernel(const float* __restrict__ energy, int R, int rep_lo, float target,
                                    int round, volatile int* hit) {
    __shared__ float s_e[32];
    __shared__ int s_r[32];
    float best = 3.0e38f;
    int arg = 0x7FFFFFFF;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const float v = energy[r];
        if (v < best) { best = v; arg = r; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float v = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        const int w = __shfl_xor_sync(0xFFFFFFFFu, arg, o);
        if (v < best || (v == best && w < arg)) { best = v; arg = w; }
    }
    if ((threadIdx.x & 31) == 0) { s_e[threadIdx.x >> 5] = best; s_r[threadIdx.x >> 5] = arg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (s_e[w] < best || (s_e[w] == best && s_r[w] < arg)) { best = s_e[w]; arg = s_r[w]; }
        if (arg != 0x7FFFFFFF && best <= target && hit[0] < 0) {
            hit[1] = rep_lo + arg;
            __threadfence_system();
            hit[0] = round;
        }
    }
}

// temperatures of the local replicas from the rung -> replica map (checkpoint restore)
__global__ void ladder_temps_kernel(const int* __restrict__ rep_at, const double* __restrict__ ladder,
                                    double* __restrict__ rep_temp, int n_global, int K, int rep_lo, int rep_n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_global) return;
    const int loc = rep_at[t] - rep_lo;
    if (loc >= 0 && loc < rep_n) rep_temp[loc] = ladder[t % K];
}

// out_idx[0] = argmin_r energy[r] (lowest index on ties), out_e[0] = that energy; then row
// out_idx[0] of the padded int8 matrix `rows` is copied to out_row (n entries).  One block.
__global__ void best_config_kernel(const float* __restrict__ energy, int R, const int8_t* __restrict__ rows,
                                   int n, int n_pad, float* out_e, int* out_idx, int8_t* out_row) {
    __shared__ float s_e[32];
    __shared__ int s_r[32];
    __shared__ int s_best;
    float best = 3.0e38f;
    int arg = 0x7FFFFFFF;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const float v = energy[r];
        if (v < best) { best = v; arg = r; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float v = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        const int w = __shfl_xor_sync(0xFFFFFFFFu, arg, o);
        if (v < best || (v == best && w < arg)) { best = v; arg = w; }
    }
    if ((threadIdx.x & 31) == 0) { s_e[threadIdx.x >> 5] = best; s_r[threadIdx.x >> 5] = arg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (s_e[w] < best || (s_e[w] == best && s_r[w] < arg)) { best = s_e[w]; arg = s_r[w]; }
        if (arg == 0x7FFFFFFF) arg = 0;
        s_best = arg;
        if (out_e) out_e[0] = best;
        if (out_idx) out_idx[0] = arg;
    }
    __syncthreads();
    if (out_row) {
        const int8_t* src = rows + (size_t)s_best * n_pad;
        for (int i = threadIdx.x; i < n; i += blockDim.x) out_row[i] = src[i];
    }
}

__global__ void ladder_init_kernel(int* rep_at, double* rep_temp, const double* ladder, int n_global,
                                   int K, int rep_lo, int rep_n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_global) return;
    rep_at[t] = t;  // replica l*K + k starts on rung k of ladder l
    const int loc = t - rep_lo;
    if (loc >= 0 && loc < rep_n) rep_temp[loc] = ladder[t % K];
}

// ---- the reference's operator form of the exchange (annealing/cuda_kernels.py:326-369 with the
// loop that actually runs, _parallel_tempering_fallback :405-436): ONE pass over the pairs
// (i, i+1), i = 0 .. R-2, in order, each using the energies left by the previous swap;
// p = exp((1/T[i+1] - 1/T[i]) * (E[i] - E[i+1])) in float32 exactly as written there (note the
// sign: it is the inverse of ParallelTempering._attempt_single_exchange's), a uniform is drawn
// for every pair, configurations AND energies are swapped, temperatures stay with the index.
// The pass is a dependency chain over R-1 scalars: one thread decides it and records where every
// row has to come from; the rows are then moved by the whole grid.
__global__ void exchange_chain_decide_kernel(float* __restrict__ energies, const float* __restrict__ temps,
                                             const float* __restrict__ uniforms, unsigned long long seed,
                                             unsigned long long round, int R, int* __restrict__ src,
                                             int* __restrict__ accepted) {
    if (threadIdx.x != 0) return;
    int acc = 0;
    // (e_cur, s_cur): energy and original index of the row that sits at position i when pair i
    // is examined -- row i itself, or the row a chain of accepted swaps has carried there
    float e_cur = energies[0];
    int s_cur = 0;
    for (int i = 0; i + 1 < R; ++i) {
        const float beta1 = 1.0f / temps[i];
        const float beta2 = 1.0f / temps[i + 1];
        const float e_next = energies[i + 1];
        const float prob = expf((beta2 - beta1) * (e_cur - e_next));
        float u;
        if (uniforms) {
            u = uniforms[i];
        } else {
            const uint4 x = philox4x32_10(
                make_uint4(0xC4A1E8C4u, (uint32_t)round, (uint32_t)(round >> 32), (uint32_t)i),
                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            u = u01(x.x);
        }
        if (u < prob) {
            // rows i and i+1 trade places: position i keeps what was at i+1, the row that was
            // at i moves on to i+1 and meets the next pair
            energies[i] = e_next;
            src[i] = i + 1;
            ++acc;
        } else {
            energies[i] = e_cur;
            src[i] = s_cur;
            e_cur = e_next;
            s_cur = i + 1;
        }
    }
    energies[R - 1] = e_cur;
    src[R - 1] = s_cur;
    *accepted = acc;
}

// dst row i <- src row map[i] (map == nullptr: identity), only rows that moved
__global__ void move_rows_kernel(unsigned char* __restrict__ dst, long long dst_stride,
                                 const unsigned char* __restrict__ src, long long src_stride,
                                 long long row_bytes, const int* __restrict__ map_to_src,
                                 const int* __restrict__ moved_ref, int vec16) {
    const int row = blockIdx.y;
    const int from = map_to_src ? map_to_src[row] : row;
    if (moved_ref[row] == row) return;
    unsigned char* d = dst + (size_t)row * dst_stride;
    const unsigned char* s = src + (size_t)from * src_stride;
    if (vec16) {
        const long long nv = row_bytes >> 4;
        for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nv;
             v += (long long)gridDim.x * blockDim.x)
            reinterpret_cast<uint4*>(d)[v] = reinterpret_cast<const uint4*>(s)[v];
    } else {
        for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < row_bytes;
             v += (long long)gridDim.x * blockDim.x)
            d[v] = s[v];
    }
}

// ---- ADAPTIVE temperature schedule evaluated on the device (TemperatureScheduler's
// AdaptiveSchedule.update, reference annealing/temperature_scheduler.py:206-249): geometric base
// temperature (precomputed by the host), multiplied by (1 -/+ adaptation_rate) once
// `window` cumulative acceptance rates of the tracked replica have been seen and their mean is
// above / below the target.  One thread, launched between two sweeps: the host never has to
// read the acceptance counter back.
//   state[0 .. window-1] = ring of the last rates, state[window] = number of rates seen
__global__ void adaptive_temperature_kernel(const unsigned long long* __restrict__ accepted,
                                            unsigned long long accepted_base, int n, int sweep,
                                            int window, double target, double rate, double t_final,
                                            const double* __restrict__ base_temps,
                                            double* __restrict__ state, double* __restrict__ temps_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double attempts = (double)sweep * (double)n;
    const double acc = (double)(accepted[0] - accepted_base);
    const double r = attempts > 0.0 ? acc / attempts : 0.0;
    int cnt = (int)state[window];
    state[cnt % window] = r;
    ++cnt;
    state[window] = (double)cnt;
    const double base = base_temps[sweep];
    double cur = base;
    if (cnt >= window) {
        double sum = 0.0;
        for (int k = 0; k < window; ++k) sum += state[(cnt + k) % window];   // oldest first
        const double recent = sum / (double)window;
        double factor = 1.0;
        if (recent > target) factor = 1.0 - rate;
        else if (recent < target) factor = 1.0 + rate;
        cur = fmax(base * factor, t_final);
    }
    temps_out[sweep] = fmax(cur, 1e-10);
}

}  // namespace

cudaError_t launch_adaptive_temperature(const unsigned long long* accepted, unsigned long long accepted_base,
                                        int n, int sweep, int window, double target, double rate,
                                        double t_final, const double* base_temps, double* state,
                                        double* temps_out, cudaStream_t st) {
    adaptive_temperature_kernel<<<1, 32, 0, st>>>(accepted, accepted_base, n, sweep, window, target, rate,
                                                 t_final, base_temps, state, temps_out);
    return cudaGetLastError();
}

cudaError_t launch_exchange_chain(void* rows, long long row_stride, long long row_bytes, int R,
                                  float* energies, const float* temps, const float* uniforms,
                                  unsigned long long seed, unsigned long long round, void* scratch,
                                  cudaStream_t st) {
    // scratch: [R] int source map, [1] int count (16-byte padded), then R * row_bytes of rows
    int* src = static_cast<int*>(scratch);
    int* acc = src + R;
    unsigned char* tmp = static_cast<unsigned char*>(scratch) + exchange_chain_header_bytes(R);
    exchange_chain_decide_kernel<<<1, 32, 0, st>>>(energies, temps, uniforms, seed, round, R, src, acc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int vec16 = (row_bytes % 16 == 0 && row_stride % 16 == 0 &&
                       reinterpret_cast<uintptr_t>(rows) % 16 == 0) ? 1 : 0;
    const long long units = vec16 ? row_bytes / 16 : row_bytes;
    int gx = (int)((units + 255) / 256);
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)R);
    move_rows_kernel<<<grid, 256, 0, st>>>(tmp, row_bytes, static_cast<const unsigned char*>(rows),
                                           row_stride, row_bytes, src, src, vec16);
    move_rows_kernel<<<grid, 256, 0, st>>>(static_cast<unsigned char*>(rows), row_stride, tmp, row_bytes,
                                           row_bytes, nullptr, src, vec16);
    return cudaGetLastError();
}

size_t exchange_chain_header_bytes(int R) { return (((size_t)R + 1) * sizeof(int) + 15) & ~(size_t)15; }

cudaError_t launch_exchange(ExchangeDev a, cudaStream_t st) {
    if (a.L <= 0 || a.K < 2) return cudaSuccess;
    if (a.method == 1) {
        exchange_all_pairs_kernel<<<(a.L + 63) / 64, 64, 0, st>>>(a);
    } else {
        if ((a.K - a.parity) / 2 <= 0) return cudaSuccess;
        exchange_kernel<<<(a.L + 3) / 4, 128, 0, st>>>(a);   // one warp per ladder
    }
    return cudaGetLastError();
}

cudaError_t launch_check_target(const float* energy, int R, int rep_lo, float target, int round, int* hit,
                                cudaStream_t st) {
    check_target_kernel<<<1, 1024, 0, st>>>(energy, R, rep_lo, target, round, hit);
    return cudaGetLastError();
}

cudaError_t launch_ladder_temps(const int* rep_at, const double* ladder, double* rep_temp, int n_global, int K,
                                int rep_lo, int rep_n, cudaStream_t st) {
    ladder_temps_kernel<<<(n_global + 127) / 128, 128, 0, st>>>(rep_at, ladder, rep_temp, n_global, K, rep_lo, rep_n);
    return cudaGetLastError();
}

cudaError_t launch_best_config(const float* energy, int R, const int8_t* rows, int n, int n_pad, float* out_e,
                               int* out_idx, int8_t* out_row, cudaStream_t st) {
    best_config_kernel<<<1, 1024, 0, st>>>(energy, R, rows, n, n_pad, out_e, out_idx, out_row);
    return cudaGetLastError();
}

cudaError_t launch_ladder_init(int* rep_at, double* rep_temp, const double* ladder, int n_global, int K,
                               int rep_lo, int rep_n, cudaStream_t st) {
    ladder_init_kernel<<<(n_global + 127) / 128, 128, 0, st>>>(rep_at, rep_temp, ladder, n_global, K,
                                                                rep_lo, rep_n);
    return cudaGetLastError();
}

}  // namespace sg
