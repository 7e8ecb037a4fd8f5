// sg_wolff.cu -- K1-WOLFF: the reference's cluster move (UpdateRule.WOLFF,
// core/spin_dynamics.py:193-262, dense branch _wolff_cluster_dense) for R replicas at once.
//
// What the reference does per update: breadth-first growth from a start site; the FIFO queue is a
// Python list, every dequeued site walks ALL n columns of its coupling row in index order with
// three .item() calls per column, and draws a uniform for each column that is not yet in the
// cluster, has coupling < 0 and the same spin; accepted with probability 1 - exp(2 J / T).  The
// whole cluster is flipped at the end; a sweep is n such updates.
//
// Two forms (launch_wolff picks): a WARP per replica for n <= 1792 (further down), and, for larger
// dense models, ONE CTA PER REPLICA with the replica's spins, the cluster bitmap and the queue in shared
// memory.  In the CTA form a dequeued site's row (n_pad floats of the row-major copy of J, L2 resident) is read
// once by the 256 threads, four consecutive columns per thread and pass, all passes in flight
// together.  Nothing in one row walk depends on another column of the same walk (a column is
// visited once, cluster membership only changes for columns accepted in this very walk), so the
// walk is evaluated in parallel and only the ORDER is restored afterwards:
//   * injected uniforms (replay of the reference's stream): an ordered block-wide prefix count of
//     the candidate columns gives every candidate the position of its uniform in the replica's
//     stream -- candidate k of the walk consumes uniform cursor + k, exactly the reference's order;
//   * a second ordered prefix count over the accepted columns appends them to the queue in index
//     order (the reference's queue.append order), so the walks that follow are the reference's.
// In Philox mode the uniform of column j in walk `visit` of update `upd` is a pure function of
// (seed, replica, upd, visit, j): no candidate count is needed, one barrier per walk that grows
// nothing.  Both counts ride on one packed warp scan (8 bits per pass) and __syncthreads_or.
#include <cstdlib>

#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

namespace {

constexpr int kWolffThreads = 256;
constexpr int kWolffWarps = kWolffThreads / 32;
constexpr int kWolffCols = kWolffThreads * 4;   // columns per pass
constexpr uint32_t kWolffKeyTag = 0x574F4C46u;  // keeps the stream apart from the single-spin rules

// probability that an aligned neighbour with coupling c < 0 joins the cluster.  Replay mode follows
// the reference's arithmetic -- 1.0 - torch.exp(torch.tensor(2.0 * coupling / T)): double quotient
// rounded to float32, exp and the subtraction in float32 (core/spin_dynamics.py:237); Philox mode
// (its own random numbers anyway) stays in float32.
template <bool INJECT>
__device__ __forceinline__ float wolff_prob(float c, double T, float inv_t) {
    if (INJECT) return 1.0f - expf((float)(2.0 * (double)c / T));
    return 1.0f - __expf(2.0f * c * inv_t);
}

// ordered block-wide ranks of up to NP x 4 flags per thread (pass-major, then thread, then bit).
// cnt[p] <= 4 per thread; a warp's sum per pass is <= 128 and fits a byte of the packed scan.
template <int NP>
struct OrderedCount {
    uint32_t before[NP];   // flags ahead of this thread's first flag of pass p (whole block order)
    uint32_t total;        // flags in the block
};

template <int NP>
__device__ __forceinline__ void packed_warp_scan(const uint32_t (&cnt)[NP], uint32_t (&incl)[2]) {
    incl[0] = incl[1] = 0u;
#pragma unroll
    for (int p = 0; p < NP; ++p) incl[p >> 2] |= cnt[p] << (8 * (p & 3));
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a0 = __shfl_up_sync(0xFFFFFFFFu, incl[0], d);
        const uint32_t a1 = (NP > 4) ? __shfl_up_sync(0xFFFFFFFFu, incl[1], d) : 0u;
        if (lane >= d) {
            incl[0] += a0;
            incl[1] += a1;
        }
    }
}

// wt: shared [kWolffWarps][2] packed warp totals, already published (barrier passed)
template <int NP>
__device__ __forceinline__ OrderedCount<NP> finish_count(const uint32_t (&cnt)[NP],
                                                         const uint32_t (&incl)[2],
                                                         const uint32_t (*wt)[2]) {
    OrderedCount<NP> oc;
    const int warp = threadIdx.x >> 5;
    uint32_t run = 0u;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int sh = 8 * (p & 3), wd = p >> 2;
        uint32_t below = 0u, all = 0u;
#pragma unroll
        for (int w = 0; w < kWolffWarps; ++w) {
            const uint32_t t = (wt[w][wd] >> sh) & 0xFFu;
            all += t;
            below += (w < warp) ? t : 0u;
        }
        const uint32_t in_warp = ((incl[wd] >> sh) & 0xFFu) - cnt[p];   // exclusive
        oc.before[p] = run + below + in_warp;
        run += all;
    }
    oc.total = run;
    return oc;
}

template <int NP, bool INJECT>
__global__ void __launch_bounds__(kWolffThreads) wolff_kernel(const WolffDev a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, n_pad = a.n_pad;
    int8_t* spin = reinterpret_cast<int8_t*>(smem);                          // [n_pad]
    uint32_t* incl_bits = reinterpret_cast<uint32_t*>(smem + n_pad);         // [n_pad / 32]
    int* queue = reinterpret_cast<int*>(smem + n_pad + (n_pad / 32) * 4);    // [n]
    __shared__ uint32_t wt_cand[kWolffWarps][2], wt_acc[kWolffWarps][2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rep = blockIdx.x;
    int8_t* spins_g = a.spins + (size_t)rep * n_pad;
    for (int i = tid; i < n_pad / 16; i += kWolffThreads)
        reinterpret_cast<int4*>(spin)[i] = reinterpret_cast<const int4*>(spins_g)[i];

    const double T = a.temps[(long long)a.sweep * a.t_ss + (long long)rep * a.t_rs];
    const float inv_t = (float)(1.0 / T);
    const int* sites = a.sites + (long long)rep * a.s_rs + (long long)a.sweep * a.s_ss;
    const float* ustream = INJECT ? a.uniforms + (long long)rep * a.u_rs : nullptr;
    long long cur = INJECT ? a.cursor[rep] : 0;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ kWolffKeyTag);
    const unsigned long long upd0 = a.sweep_abs * (unsigned long long)n;
    unsigned long long flips = 0;
    bool dry = false;

#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const int start = sites[k];
        for (int w = tid; w < n_pad / 32; w += kWolffThreads) incl_bits[w] = 0u;
        __syncthreads();   // spins loaded / previous flips done, bitmap clear
        if (tid == 0) {
            queue[0] = start;
            incl_bits[start >> 5] = 1u << (start & 31);
        }
        __syncthreads();
        int head = 0, tail = 1;
        const unsigned long long upd = upd0 + (unsigned long long)k;
#pragma unroll 1
        while (head < tail) {
            const int c = queue[head];
            const int8_t sc = spin[c];
            const float4* row = reinterpret_cast<const float4*>(a.Jrow + (size_t)c * n_pad);
            float4 v[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int j4 = p * kWolffCols + tid * 4;
                v[p] = (j4 < n_pad) ? __ldg(row + (j4 >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // candidates: not the site itself, not in the cluster, coupling < 0, same spin
            uint32_t cand[NP], cnt[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int j4 = p * kWolffCols + tid * 4;
                uint32_t m = 0u;
                if (j4 < n) {
                    const uint32_t inb = (incl_bits[j4 >> 5] >> (j4 & 31)) & 0xFu;
                    const uint32_t sp = *reinterpret_cast<const uint32_t*>(spin + j4);
                    const float jv[4] = {v[p].x, v[p].y, v[p].z, v[p].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = j4 + i;
                        const bool same = (int8_t)((sp >> (8 * i)) & 0xFFu) == sc;
                        if (j < n && j != c && !((inb >> i) & 1u) && jv[i] < 0.0f && same) m |= 1u << i;
                    }
                }
                cand[p] = m;
                cnt[p] = __popc(m);
            }
            uint32_t mine = 0u;
#pragma unroll
            for (int p = 0; p < NP; ++p) mine |= cand[p];

            uint32_t n_cand = 0u;
            OrderedCount<NP> oc_c;
            if (INJECT) {
                uint32_t inc[2];
                packed_warp_scan<NP>(cnt, inc);
                if (lane == 31) {
                    wt_cand[warp][0] = inc[0];
                    wt_cand[warp][1] = inc[1];
                }
                if (!__syncthreads_or(mine != 0u)) {   // nobody to ask: the walk draws nothing
                    ++head;
                    continue;
                }
                oc_c = finish_count<NP>(cnt, inc, wt_cand);
                n_cand = oc_c.total;
            }
            // decisions
            uint32_t acc[NP], acnt[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                uint32_t m = 0u;
                if (cand[p]) {
                    const int j4 = p * kWolffCols + tid * 4;
                    const float jv[4] = {v[p].x, v[p].y, v[p].z, v[p].w};
                    float u4[4];
                    if (!INJECT) {
                        const uint4 x = philox4x32_10(
                            make_uint4((uint32_t)(rep + a.rep_base), (uint32_t)upd, (uint32_t)(upd >> 32),
                                       ((uint32_t)head << 11) | (uint32_t)(j4 >> 2)), key);
                        u4[0] = u01(x.x); u4[1] = u01(x.y); u4[2] = u01(x.z); u4[3] = u01(x.w);
                    }
                    uint32_t rank = INJECT ? oc_c.before[p] : 0u;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (!((cand[p] >> i) & 1u)) continue;
                        float u;
                        if (INJECT) {
                            const long long at = cur + (long long)rank;
                            ++rank;
                            if (at < a.u_len) u = ustream[at];
                            else { u = 2.0f; dry = true; }
                        } else {
                            u = u4[i];
                        }
                        if (u < wolff_prob<INJECT>(jv[i], T, inv_t)) m |= 1u << i;
                    }
                }
                acc[p] = m;
                acnt[p] = __popc(m);
            }
            uint32_t grown = 0u;
#pragma unroll
            for (int p = 0; p < NP; ++p) grown |= acc[p];
            uint32_t inc2[2];
            packed_warp_scan<NP>(acnt, inc2);
            if (lane == 31) {
                wt_acc[warp][0] = inc2[0];
                wt_acc[warp][1] = inc2[1];
            }
            cur += (long long)n_cand;
            ++head;
            if (!__syncthreads_or(grown != 0u)) continue;   // the cluster did not grow in this walk
            const OrderedCount<NP> oc_a = finish_count<NP>(acnt, inc2, wt_acc);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (!acc[p]) continue;
                const int j4 = p * kWolffCols + tid * 4;
                int at = tail + (int)oc_a.before[p];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if ((acc[p] >> i) & 1u) queue[at++] = j4 + i;
                atomicOr(&incl_bits[j4 >> 5], acc[p] << (j4 & 31));
            }
            tail += (int)oc_a.total;
            __syncthreads();   // queue and bitmap of this walk visible to the next one
        }
        for (int q = tid; q < tail; q += kWolffThreads) {
            const int j = queue[q];
            spin[j] = (int8_t)-spin[j];
        }
        flips += (unsigned long long)tail;
    }
    __syncthreads();
    for (int i = tid; i < n_pad / 16; i += kWolffThreads)
        reinterpret_cast<int4*>(spins_g)[i] = reinterpret_cast<const int4*>(spin)[i];
    if (tid == 0) {
        a.accepted[rep] += flips;
        if (INJECT) a.cursor[rep] = cur;
    }
    if (INJECT && dry) *a.status = 1;
}

// ---------------------------------------------------------------------------------------------
// n <= 1792: ONE WARP PER REPLICA (eight replicas per CTA).  A row walk of a small model is a
// latency chain (one L2 round trip, then the two ordered counts); a whole CTA per replica leaves
// the SM with four walks in flight.  Here a warp reads the row with coalesced float4 loads (pass q,
// lane l: columns 128 q + 4 l ..), spins and cluster membership are bit masks in the registers of
// the lane that owns the column, the candidates of a walk are three ANDs, and index order (pass,
// lane, bit) is restored with a byte-per-pass shuffle scan: the walk needs no block barrier and no
// shared memory but the queue, and an SM keeps up to 32 walks in flight.  Same Philox counters per (update, visit, column quad) as the CTA form: both
// forms give the same spins in both RNG modes.
// Ordered counts of a warp's flags in index order (pass q, lane, bit): per-pass counts ride as bytes
// of one or two 64-bit words through a five-step shuffle scan (a pass holds at most 128 flags).
template <int W4>
struct PassCount {
    unsigned long long inc[2], tot[2];   // inclusive over lanes / warp totals, a byte per pass
    __device__ __forceinline__ void scan(const unsigned long long (&pk)[2], int lane) {
        inc[0] = pk[0];
        inc[1] = pk[1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t0 = __shfl_up_sync(0xFFFFFFFFu, inc[0], d);
            const unsigned long long t1 = (W4 > 8) ? __shfl_up_sync(0xFFFFFFFFu, inc[1], d) : 0ull;
            if (lane >= d) {
                inc[0] += t0;
                inc[1] += t1;
            }
        }
        tot[0] = __shfl_sync(0xFFFFFFFFu, inc[0], 31);
        tot[1] = (W4 > 8) ? __shfl_sync(0xFFFFFFFFu, inc[1], 31) : 0ull;
    }
    static __device__ __forceinline__ uint32_t byte_of(const unsigned long long (&w)[2], int q) {
        const unsigned long long x = (W4 > 8 && q >= 8) ? w[1] : w[0];   // (no dynamic indexing: registers)
        return (uint32_t)((x >> (8 * (q & 7))) & 0xFFull);
    }
    __device__ __forceinline__ uint32_t total() const {
        uint32_t t = 0u;
#pragma unroll
        for (int q = 0; q < W4; ++q) t += byte_of(tot, q);
        return t;
    }
};

// walks one lane's flags in ascending order and returns each flag's position in the warp's order
template <int W4>
struct PassCursor {
    int qc = 0, qlast = -1;
    uint32_t base = 0u, within = 0u;
    __device__ __forceinline__ uint32_t next(const PassCount<W4>& pc, const unsigned long long (&pk)[2], int q) {
        while (qc < q) {
            base += PassCount<W4>::byte_of(pc.tot, qc);
            ++qc;
        }
        if (q != qlast) {
            within = 0u;
            qlast = q;
        }
        return base + PassCount<W4>::byte_of(pc.inc, q) - PassCount<W4>::byte_of(pk, q) + within++;
    }
};

template <int W4, bool INJECT>   // W4 = float4 loads per lane and walk: 128 W4 >= n columns
__global__ void __launch_bounds__(kWolffThreads, W4 <= 8 ? 4 : 3) wolff_warp_kernel(const WolffDev a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, n_pad = a.n_pad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rep = blockIdx.x * kWolffWarps + warp;
    if (rep >= a.R) return;   // (no block-wide barrier below)
    unsigned short* queue = reinterpret_cast<unsigned short*>(smem) + (size_t)warp * ((n + 7) & ~7);

    // The replica's state lives in registers: lane l owns the columns 128 q + 4 l + i (the ones its
    // float4 loads cover), bit 4 q + i of `sb` = spin up, of `ic` = in the cluster.  A lane only
    // ever appends its own columns, so membership needs no shared bitmap, and flipping the cluster
    // at the end of an update is sb ^= ic.
    int8_t* spins_g = a.spins + (size_t)rep * n_pad;
    unsigned long long sb = 0ull;
#pragma unroll
    for (int q = 0; q < W4; ++q) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(spins_g + 128 * q + 4 * lane);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if ((int8_t)((w >> (8 * i)) & 0xFFu) > 0) sb |= 1ull << (4 * q + i);
    }
    const double T = a.temps[(long long)a.sweep * a.t_ss + (long long)rep * a.t_rs];
    const float inv_t = (float)(1.0 / T);
    const int* sites = a.sites + (long long)rep * a.s_rs + (long long)a.sweep * a.s_ss;
    const float* ustream = INJECT ? a.uniforms + (long long)rep * a.u_rs : nullptr;
    long long cur = INJECT ? a.cursor[rep] : 0;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ kWolffKeyTag);
    const unsigned long long upd0 = a.sweep_abs * (unsigned long long)n;
    unsigned long long flips = 0;
    bool dry = false;

#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const int start = sites[k];
        unsigned long long ic = 0ull;
        if (((start & 127) >> 2) == lane) ic = 1ull << (4 * (start >> 7) + (start & 3));
        if (lane == 0) queue[0] = (unsigned short)start;
        __syncwarp();
        int head = 0, tail = 1;
        const unsigned long long upd = upd0 + (unsigned long long)k;
#pragma unroll 1
        while (head < tail) {
            const int c = queue[head];
            const bool c_up = (__shfl_sync(0xFFFFFFFFu, sb, (c & 127) >> 2) >> (4 * (c >> 7) + (c & 3))) & 1ull;
            const float* rowf = a.Jrow + (size_t)c * n_pad;
            // candidates: coupling < 0 (padding columns hold 0), same spin, not in the cluster (the
            // dequeued site itself is).  The couplings are only needed again for the few candidates:
            // re-read through L1 below, so the row leaves the registers here.
            unsigned long long nm = 0ull;
            {
                const float4* row = reinterpret_cast<const float4*>(rowf) + lane;
                float4 v[W4];
#pragma unroll
                for (int q = 0; q < W4; ++q) v[q] = __ldg(row + 32 * q);
#pragma unroll
                for (int q = 0; q < W4; ++q) {
                    const uint32_t m = (v[q].x < 0.0f ? 1u : 0u) | (v[q].y < 0.0f ? 2u : 0u) |
                                       (v[q].z < 0.0f ? 4u : 0u) | (v[q].w < 0.0f ? 8u : 0u);
                    nm |= (unsigned long long)m << (4 * q);
                }
            }
            const unsigned long long cm = nm & (c_up ? sb : ~sb) & ~ic;
            const int visit = head++;
            if (!__any_sync(0xFFFFFFFFu, cm != 0ull)) continue;   // the walk draws nothing
            PassCount<W4> pc;
            unsigned long long pk[2] = {0ull, 0ull};   // flags per pass, a byte each
            if (INJECT) {   // index order = pass, lane, bit
#pragma unroll
                for (int q = 0; q < W4; ++q)
                    pk[q >> 3] += (unsigned long long)__popc((uint32_t)(cm >> (4 * q)) & 0xFu) << (8 * (q & 7));
                pc.scan(pk, lane);
            }
            unsigned long long am = 0ull;
            {
                PassCursor<W4> pos;
#pragma unroll 1
                for (unsigned long long rest = cm; rest; rest &= rest - 1) {
                    const int b = __ffsll((long long)rest) - 1;
                    const int q = b >> 2, i = b & 3;
                    const int col = 128 * q + 4 * lane + i;
                    float u;
                    if (INJECT) {
                        const long long at = cur + (long long)pos.next(pc, pk, q);
                        if (at < a.u_len) u = ustream[at];
                        else { u = 2.0f; dry = true; }
                    } else {   // the uniform of a column is a function of (update, visit, column quad)
                        const uint4 x = philox4x32_10(
                            make_uint4((uint32_t)(rep + a.rep_base), (uint32_t)upd, (uint32_t)(upd >> 32),
                                       ((uint32_t)visit << 11) | (uint32_t)(col >> 2)), key);
                        u = u01(i == 0 ? x.x : i == 1 ? x.y : i == 2 ? x.z : x.w);
                    }
                    if (u < wolff_prob<INJECT>(__ldg(rowf + col), T, inv_t)) am |= 1ull << b;
                }
            }
            if (INJECT) cur += (long long)pc.total();
            if (!__any_sync(0xFFFFFFFFu, am != 0ull)) continue;   // the cluster did not grow
#pragma unroll
            for (int q = 0; q < 2; ++q) pk[q] = 0ull;
#pragma unroll
            for (int q = 0; q < W4; ++q)
                pk[q >> 3] += (unsigned long long)__popc((uint32_t)(am >> (4 * q)) & 0xFu) << (8 * (q & 7));
            pc.scan(pk, lane);
            {
                PassCursor<W4> pos;
#pragma unroll 1
                for (unsigned long long rest = am; rest; rest &= rest - 1) {
                    const int b = __ffsll((long long)rest) - 1;
                    const int q = b >> 2;
                    queue[tail + (int)pos.next(pc, pk, q)] = (unsigned short)(128 * q + 4 * lane + (b & 3));
                }
            }
            ic |= am;
            tail += (int)pc.total();
            __syncwarp();   // queue entries of this walk visible to the next one
        }
        sb ^= ic;   // flip the cluster
        flips += (unsigned long long)tail;
        __syncwarp();   // every lane is done with the queue before the next update reuses it
    }
#pragma unroll
    for (int q = 0; q < W4; ++q) {
        uint32_t w = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) w |= (((sb >> (4 * q + i)) & 1ull) ? 0x01u : 0xFFu) << (8 * i);
        *reinterpret_cast<uint32_t*>(spins_g + 128 * q + 4 * lane) = w;
    }
    if (lane == 0) {
        a.accepted[rep] += flips;
        if (INJECT) a.cursor[rep] = cur;
    }
    if (INJECT && dry) *a.status = 1;
}

// ---------------------------------------------------------------------------------------------
// Models whose rows hold at most 32 NEGATIVE couplings (lattices, sparse graphs stored dense): only
// those columns can ever be candidates, so a walk does not need the row at all.  The engine builds,
// once per model, the ascending list of negative columns of every row (32 slots per row: column,
// value); a walk is ONE LANE PER LIST ENTRY -- index order is lane order, the two ordered counts are
// two ballots -- and touches 32 x 6 bytes instead of 4 n.  One warp per replica, any n the dense
// engine takes; spins and cluster membership are bit arrays in shared memory (cleared / flipped per
// cluster member, not per column).  Same Philox counters per (update, visit, column quad) as the
// row-walking forms: all three give the same spins in both RNG modes.
template <bool INJECT>
__global__ void __launch_bounds__(kWolffThreads) wolff_list_kernel(const WolffDev a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, n_pad = a.n_pad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rep = blockIdx.x * kWolffWarps + warp;
    if (rep >= a.R) return;   // (no block-wide barrier below)
    const int words = (n + 31) / 32;
    const size_t per_warp = (((size_t)n * 2 + (size_t)words * 8) + 15) & ~(size_t)15;
    unsigned char* base = smem + (size_t)warp * per_warp;
    uint32_t* sbits = reinterpret_cast<uint32_t*>(base);              // [words] spin up
    uint32_t* cbits = sbits + words;                                   // [words] in the cluster
    unsigned short* queue = reinterpret_cast<unsigned short*>(cbits + words);   // [n]

    int8_t* spins_g = a.spins + (size_t)rep * n_pad;
    for (int w = lane; w < words; w += 32) {
        uint32_t m = 0u;
        for (int i = 0; i < 32; ++i) {
            const int j = w * 32 + i;
            if (j < n && spins_g[j] > 0) m |= 1u << i;
        }
        sbits[w] = m;
        cbits[w] = 0u;
    }
    const double T = a.temps[(long long)a.sweep * a.t_ss + (long long)rep * a.t_rs];
    const float inv_t = (float)(1.0 / T);
    const int* sites = a.sites + (long long)rep * a.s_rs + (long long)a.sweep * a.s_ss;
    const float* ustream = INJECT ? a.uniforms + (long long)rep * a.u_rs : nullptr;
    long long cur = INJECT ? a.cursor[rep] : 0;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ kWolffKeyTag);
    const unsigned long long upd0 = a.sweep_abs * (unsigned long long)n;
    unsigned long long flips = 0;
    bool dry = false;
    const uint32_t lt = (1u << lane) - 1u;
    __syncwarp();

#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const int start = sites[k];
        if (lane == 0) {
            queue[0] = (unsigned short)start;
            cbits[start >> 5] = 1u << (start & 31);   // (all other cluster bits are clear here)
        }
        __syncwarp();
        int head = 0, tail = 1;
        const unsigned long long upd = upd0 + (unsigned long long)k;
#pragma unroll 1
        while (head < tail) {
            const int c = queue[head];
            const bool c_up = (sbits[c >> 5] >> (c & 31)) & 1u;
            const int j = a.nb_col[(size_t)c * 32 + lane];          // ascending; 0xFFFF beyond the row's list
            bool cand = false;
            if (j < n) {
                const bool up = (sbits[j >> 5] >> (j & 31)) & 1u;
                const bool in = (cbits[j >> 5] >> (j & 31)) & 1u;
                cand = (up == c_up) && !in;                           // (j == c is in the cluster)
            }
            const uint32_t cmask = __ballot_sync(0xFFFFFFFFu, cand);
            const int visit = head++;
            if (cmask == 0u) continue;                                // the walk draws nothing
            bool take = false;
            if (cand) {
                float u;
                if (INJECT) {
                    const long long at = cur + (long long)__popc(cmask & lt);
                    if (at < a.u_len) u = ustream[at];
                    else { u = 2.0f; dry = true; }
                } else {
                    const uint4 x = philox4x32_10(
                        make_uint4((uint32_t)(rep + a.rep_base), (uint32_t)upd, (uint32_t)(upd >> 32),
                                   ((uint32_t)visit << 11) | (uint32_t)(j >> 2)), key);
                    const int i = j & 3;
                    u = u01(i == 0 ? x.x : i == 1 ? x.y : i == 2 ? x.z : x.w);
                }
                take = u < wolff_prob<INJECT>(a.nb_val[(size_t)c * 32 + lane], T, inv_t);
            }
            if (INJECT) cur += (long long)__popc(cmask);
            const uint32_t amask = __ballot_sync(0xFFFFFFFFu, take);
            if (amask == 0u) continue;                                // the cluster did not grow
            if (take) {
                queue[tail + __popc(amask & lt)] = (unsigned short)j;
                atomicOr(&cbits[j >> 5], 1u << (j & 31));
            }
            tail += __popc(amask);
            __syncwarp();   // queue and bit arrays of this walk visible to the next one
        }
        for (int q = lane; q < tail; q += 32) {   // flip the cluster, take its members out again
            const int j = queue[q];
            atomicXor(&sbits[j >> 5], 1u << (j & 31));
            atomicAnd(&cbits[j >> 5], ~(1u << (j & 31)));
        }
        flips += (unsigned long long)tail;
        __syncwarp();
    }
    for (int j = lane; j < n; j += 32) spins_g[j] = ((sbits[j >> 5] >> (j & 31)) & 1u) ? (int8_t)1 : (int8_t)-1;
    if (lane == 0) {
        a.accepted[rep] += flips;
        if (INJECT) a.cursor[rep] = cur;
    }
    if (INJECT && dry) *a.status = 1;
}

// number of negative couplings of every row and their maximum (one warp per row)
__global__ void wolff_neg_count_kernel(const float* __restrict__ Jrow, int n, int n_pad, int* __restrict__ max_deg) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    int cnt = 0;
    for (int j = lane; j < n; j += 32) cnt += Jrow[(size_t)row * n_pad + j] < 0.0f;
    for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
    if (lane == 0) atomicMax(max_deg, cnt);
}

// the ascending list of negative columns of every row, 32 slots per row (only called when they fit)
__global__ void wolff_neg_fill_kernel(const float* __restrict__ Jrow, int n, int n_pad,
                                      unsigned short* __restrict__ nb_col, float* __restrict__ nb_val) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    nb_col[(size_t)row * 32 + lane] = 0xFFFFu;
    nb_val[(size_t)row * 32 + lane] = 0.0f;
    __syncwarp();
    int at = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const float v = (j < n) ? Jrow[(size_t)row * n_pad + j] : 0.0f;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, v < 0.0f);
        if (v < 0.0f) {
            const int slot = at + __popc(m & ((1u << lane) - 1u));
            if (slot < 32) {
                nb_col[(size_t)row * 32 + slot] = (unsigned short)j;
                nb_val[(size_t)row * 32 + slot] = v;
            }
        }
        at += __popc(m);
    }
}

// after the exact energy refresh of a sweep: trace row and compare-and-keep best
__global__ void wolff_record_kernel(const float* __restrict__ energy, float* __restrict__ best_energy,
                                    const int8_t* __restrict__ spins, int8_t* __restrict__ best_spins,
                                    float* __restrict__ trace_row, int n_pad, int track_best) {
    const int rep = blockIdx.x;
    const float e = energy[rep];
    if (trace_row && threadIdx.x == 0) trace_row[rep] = e;
    if (!track_best) return;
    const bool better = e < best_energy[rep];   // every thread reads before anyone writes: see barrier
    __syncthreads();
    if (!better) return;
    if (threadIdx.x == 0) best_energy[rep] = e;
    const int4* src = reinterpret_cast<const int4*>(spins + (size_t)rep * n_pad);
    int4* dst = reinterpret_cast<int4*>(best_spins + (size_t)rep * n_pad);
    for (int i = threadIdx.x; i < n_pad / 16; i += blockDim.x) dst[i] = src[i];
}

template <int NP>
cudaError_t launch_np(const WolffDev& a, bool inject, size_t smem, cudaStream_t st) {
    if (inject)
        wolff_kernel<NP, true><<<a.R, kWolffThreads, smem, st>>>(a);
    else
        wolff_kernel<NP, false><<<a.R, kWolffThreads, smem, st>>>(a);
    return cudaGetLastError();
}

template <int W4>
cudaError_t launch_warp(const WolffDev& a, bool inject, cudaStream_t st) {
    const size_t smem = (size_t)kWolffWarps * ((a.n + 7) & ~7) * sizeof(unsigned short);   // the queues
    const int grid = (a.R + kWolffWarps - 1) / kWolffWarps;
    cudaError_t e = inject ? cudaFuncSetAttribute(wolff_warp_kernel<W4, true>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                           : cudaFuncSetAttribute(wolff_warp_kernel<W4, false>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (inject)
        wolff_warp_kernel<W4, true><<<grid, kWolffThreads, smem, st>>>(a);
    else
        wolff_warp_kernel<W4, false><<<grid, kWolffThreads, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

// SG_WOLFF_FORM (tests, A/B timing): "cta" forces the CTA-per-replica form for every size, "row" the
// row-walking forms (warp or CTA by size) also for models that have neighbour lists
static bool wolff_force_cta() {
    const char* f = getenv("SG_WOLFF_FORM");
    return f && f[0] == 'c';
}
static bool wolff_force_row() {
    const char* f = getenv("SG_WOLFF_FORM");
    return f && (f[0] == 'c' || f[0] == 'r');
}

cudaError_t launch_wolff_neg_count(const float* Jrow, int n, int n_pad, int* max_deg, cudaStream_t st) {
    wolff_neg_count_kernel<<<(n + 7) / 8, 256, 0, st>>>(Jrow, n, n_pad, max_deg);
    return cudaGetLastError();
}

cudaError_t launch_wolff_neg_fill(const float* Jrow, int n, int n_pad, unsigned short* nb_col, float* nb_val,
                                  cudaStream_t st) {
    wolff_neg_fill_kernel<<<(n + 7) / 8, 256, 0, st>>>(Jrow, n, n_pad, nb_col, nb_val);
    return cudaGetLastError();
}

cudaError_t launch_wolff(const WolffDev& a, bool inject, cudaStream_t st) {
    if (a.nb_col && !wolff_force_row()) {   // at most 32 negative couplings per row: neighbour lists
        const size_t per_warp = (((size_t)a.n * 2 + (size_t)((a.n + 31) / 32) * 8) + 15) & ~(size_t)15;
        const size_t smem = per_warp * kWolffWarps;
        const int grid = (a.R + kWolffWarps - 1) / kWolffWarps;
        cudaError_t e = inject ? cudaFuncSetAttribute(wolff_list_kernel<true>,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                               : cudaFuncSetAttribute(wolff_list_kernel<false>,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (inject)
            wolff_list_kernel<true><<<grid, kWolffThreads, smem, st>>>(a);
        else
            wolff_list_kernel<false><<<grid, kWolffThreads, smem, st>>>(a);
        return cudaGetLastError();
    }
    if (!wolff_force_cta() && a.n_pad <= 1792) {   // small models: a warp per replica
        const int q = (a.n + 127) / 128;            // float4 loads per lane that cover the n columns
        if (q <= 2) return launch_warp<2>(a, inject, st);
        if (q <= 4) return launch_warp<4>(a, inject, st);
        if (q <= 7) return launch_warp<7>(a, inject, st);
        if (q <= 8) return launch_warp<8>(a, inject, st);
        if (q <= 10) return launch_warp<10>(a, inject, st);
        return launch_warp<14>(a, inject, st);
    }
    const size_t smem = (size_t)a.n_pad + (size_t)(a.n_pad / 32) * 4 + (size_t)a.n * 4;
    const int np = (a.n_pad + kWolffCols - 1) / kWolffCols;
    switch (np) {
        case 1: return launch_np<1>(a, inject, smem, st);
        case 2: return launch_np<2>(a, inject, smem, st);
        case 3: return launch_np<3>(a, inject, smem, st);
        case 4: return launch_np<4>(a, inject, smem, st);
        case 5: return launch_np<5>(a, inject, smem, st);
        case 6: return launch_np<6>(a, inject, smem, st);
        case 7: return launch_np<7>(a, inject, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_wolff_record(const float* energy, float* best_energy, const int8_t* spins,
                                int8_t* best_spins, float* trace_row, int n_pad, int R, int track_best,
                                cudaStream_t st) {
    wolff_record_kernel<<<R, 128, 0, st>>>(energy, best_energy, spins, best_spins, trace_row, n_pad,
                                           track_best);
    return cudaGetLastError();
}

}  // namespace sg
