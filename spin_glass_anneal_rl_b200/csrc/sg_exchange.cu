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
// One thread per (ladder, pair); pairs of one parity are disjoint.
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

namespace {

__global__ void exchange_kernel(const ExchangeDev a) {
    const int npairs = (a.K - a.parity) / 2;  // pairs (k, k+1), k = parity, parity+2, ... < K-1
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.L * npairs) return;
    const int l = t / npairs, m = t - l * npairs;
    const int k = a.parity + 2 * m;
    if (k + 1 >= a.K) return;
    int* slot = a.rep_at + (size_t)l * a.K;
    const int ra = slot[k], rb = slot[k + 1];
    const double Ti = a.ladder[k], Tj = a.ladder[k + 1];
    const double bi = 1.0 / Ti, bj = 1.0 / Tj;
    const double Ei = (double)a.energy[ra], Ej = (double)a.energy[rb];
    const double arg = (bj - bi) * (Ej - Ei);
    const double prob = fmin(1.0, exp(arg));
    double u;
    if (a.inject) {
        u = a.uniforms[(size_t)l * (a.K / 2) + m];
    } else {
        const uint4 x = philox4x32_10(
            make_uint4(0xE8C4A46Eu, (uint32_t)a.round, (uint32_t)(a.round >> 32), (uint32_t)t),
            make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
        u = ((double)(x.x >> 5) * 67108864.0 + (double)(x.y >> 6)) * (1.0 / 9007199254740992.0);
    }
    const size_t st = (size_t)l * (a.K - 1) + k;
    a.attempts[st] += 1u;
    if (u < prob) {
        slot[k] = rb;
        slot[k + 1] = ra;
        a.rep_temp[ra] = Tj;
        a.rep_temp[rb] = Ti;
        a.accepts[st] += 1u;
    }
}

__global__ void ladder_init_kernel(int* rep_at, double* rep_temp, const double* ladder, int L,
                                   int K) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L * K) return;
    rep_at[t] = t;  // replica l*K + k starts on rung k of ladder l
    rep_temp[t] = ladder[t % K];
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
    const int npairs = (a.K - a.parity) / 2;
    const int total = a.L * npairs;
    if (total <= 0) return cudaSuccess;
    exchange_kernel<<<(total + 127) / 128, 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_ladder_init(int* rep_at, double* rep_temp, const double* ladder, int L, int K,
                               cudaStream_t st) {
    ladder_init_kernel<<<(L * K + 127) / 128, 128, 0, st>>>(rep_at, rep_temp, ladder, L, K);
    return cudaGetLastError();
}

}  // namespace sg
