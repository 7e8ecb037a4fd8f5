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

}  // namespace

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
