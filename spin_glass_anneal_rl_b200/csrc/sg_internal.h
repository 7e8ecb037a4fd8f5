// sg_internal.h -- launcher prototypes shared by the translation units of libsg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace sg {

constexpr int kSweepThreads = 256;              // 8 warps: 2 per SM sub-partition, 255 regs each
constexpr int kBulkThreads = 224;               // 7 warps own local-field columns, 1 warp decides
constexpr int kColQuantum = 4 * kBulkThreads;   // n_pad is a multiple of this (896)

// Kernel argument block of the sweep kernel (passed by value).
struct SweepDev {
    const float* Jt;          // [n][n_pad]  row i = couplings INTO every j from i:  Jt[i][j] = J[j][i]
    const float* h;           // [n_pad]
    int8_t* spins;            // [R][n_pad]
    float* fields;            // [R][n_pad]
    float* energy;            // [R]
    float* best_energy;       // [R]
    int8_t* best_spins;       // [R][n_pad]
    unsigned long long* accepted;  // [R]
    float* energy_trace;      // [n_sweeps][R] or null
    const double* temps;      // T(s, r) = temps[s*t_ss + r*t_rs]
    long long t_ss, t_rs;
    const int* sites;         // explicit site lists
    long long s_bs, s_ss;
    const float* uniforms;    // injected uniforms [R][n_sweeps][n]
    unsigned long long seed, sweep_base;
    int n, n_pad, R, G, n_sweeps, rule, site_mode, track_best, D;
    long long* dbg;           // optional timeline buffer (development aid), block 0 only
    float* site_de;           // optional [R][n]: sum of the accepted energy changes per site
    int rpm;                  // stacked models (K1-SMALL): replicas per model, 0 = one model
    int rep_base;             // global id of replica 0 (Philox key of replica r = rep_base + r; sharded runs)
    int trace_ld;             // row length of energy_trace (0 = R): a launch over a slice of the replicas
};

// Largest number of replicas one block can hold for this padded size (0 = unsupported).
int sweep_max_replicas_per_block(int n_pad);
// Dynamic shared memory the sweep kernel needs for (n_pad, G, D).
size_t sweep_smem_bytes(int n_pad, int g_template, int D);
// Picks D, sets the smem attribute, launches.  Returns cudaError_t.
cudaError_t launch_sweep(SweepDev a, bool inject, int grid, cudaStream_t st);

// K1-SMALL (sg_sweep_small.cu): n <= 224, J in shared memory, one warp per replica.
// sites: int32 table [n_sweeps][n] from launch_sites_table.
bool sweep_small_supported(int n);
cudaError_t launch_sweep_small(const SweepDev& a, bool inject, const int* sites, cudaStream_t st);
// J [M][n][n], h [M][n] (device) -> padded column-major stacks Jt [M][n][n_pad], h_pad [M][n_pad]
cudaError_t launch_stack_models(const float* J, const float* h, int M, int n, int n_pad, float* Jt,
                                float* h_pad, cudaStream_t st);
// fields / energies of B configurations of stacked models (configuration b uses model b / rpm)
cudaError_t launch_fields_small(const float* Jt, const float* h, const int8_t* spins, float* fields,
                                float* energy, int n, int n_pad, int B, int rpm, cudaStream_t st);

// K2 + helpers (sg_fields.cu)
cudaError_t launch_pad_transpose(const float* J, int64_t ldJ, int n, float* Jt, int n_pad,
                                 cudaStream_t st);
cudaError_t launch_fields(const int8_t* spins, int64_t ld_spins, const float* Jt, const float* h,
                          int n, int n_pad, int R, float* fields, int64_t ld_fields,
                          cudaStream_t st);
cudaError_t launch_energies(const int8_t* spins, int64_t ld_spins, const float* fields,
                            int64_t ld_fields, const float* h, int n, int R, float* energy,
                            cudaStream_t st);
cudaError_t launch_pad_spins(const int8_t* src, int n, int8_t* dst, int n_pad, int R,
                             cudaStream_t st);
cudaError_t launch_unpad_spins(const int8_t* src, int n_pad, int8_t* dst, int n, int R,
                               cudaStream_t st);
cudaError_t launch_unpad_f32(const float* src, int n_pad, float* dst, int n, int R,
                             cudaStream_t st);
cudaError_t launch_stream_probe(const float4* buf, int64_t n_vec, int iters, int stagger,
                                float* sink, int grid, cudaStream_t st);

cudaError_t launch_tma_probe(const float* buf, int64_t buf_rows, uint32_t row_bytes, int n_rows,
                             int depth, int stagger, float* sink, int grid, cudaStream_t st);

// Optional per-kernel device timing (CUDA events on the launch stream): class 0 = sweep kernel,
// class 1 = operand gather.  collect() synchronises the events and returns the sums.
struct KernelTimer {
    struct Span { cudaEvent_t a, b; int cls; };
    std::vector<Span> spans;
    void begin(int cls, cudaStream_t st) {
        Span s{};
        s.cls = cls;
        cudaEventCreate(&s.a);
        cudaEventCreate(&s.b);
        cudaEventRecord(s.a, st);
        spans.push_back(s);
    }
    void end(cudaStream_t st) { cudaEventRecord(spans.back().b, st); }
    void collect(double ms[2], unsigned long long count[2]) {
        ms[0] = ms[1] = 0.0;
        count[0] = count[1] = 0;
        for (Span& s : spans) {
            float t = 0.0f;
            cudaEventSynchronize(s.b);
            cudaEventElapsedTime(&t, s.a, s.b);
            ms[s.cls] += t;
            count[s.cls]++;
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
        spans.clear();
    }
};

// K1-TC (sg_sweep_tc.cu): bf16 coupling planes and the tensor-core sweep
cudaError_t launch_split_planes(const float* Jt, int n, int n_pad, void* Jp, int n_tc, int* used,
                                cudaStream_t st);
cudaError_t launch_tc_selftest(const void* Jp, int n, int n_tc, int planes, const int* sites,
                               const float* deltas, const float* fields_in, float* fields_out,
                               cudaStream_t st);

cudaError_t launch_tc_mma_bench(int variant, int n_dim, int iters, long long* out, cudaStream_t st);
bool sweep_tc_supported(int n, int n_tc);
size_t sweep_tc_sites_bytes(int n, int n_sweeps, int R);
int sweep_tc_cluster_size(int n_tc, int R);
int sweep_tc_side_replicas(int n, int n_tc, int planes, int R, int n_sweeps);
size_t sweep_tc_stream_bytes_per_sweep(int n, int n_tc, int planes);
// sites_buf: device scratch of sweep_tc_sites_bytes(); stream_buf: device scratch for the operand
// stream, at least one sweep's worth (the launch is cut into sub-launches of as many sweeps as
// fit).  Launches the site-table kernel, then (gather, sweep) per sub-launch; counts them.
cudaError_t launch_sweep_tc(const SweepDev& a, const void* Jp, int n_tc, int planes, bool inject,
                            void* sites_buf, void* stream_buf, size_t stream_cap,
                            uint64_t* launches, KernelTimer* timer, cudaStream_t st);

// K2-TC (sg_fields_tc.cu): exact field initialisation on the int8 tensor cores
size_t fields_tc_digits_bytes(int n, int n_tc);
size_t fields_tc_spin_tiles_bytes(int n, int R);
cudaError_t launch_fields_tc_prepare(const float* Jt, int n, int n_pad, int n_tc, unsigned int* info,
                                     double* scale, void* dig, cudaStream_t st);
cudaError_t launch_fields_tc(const int8_t* spins, int64_t ld_spins, const void* dig,
                             const double* scale, const float* h, int n, int n_tc, int R,
                             void* spin_tiles, float* fields, int64_t ld_fields, cudaStream_t st);

// K1-CSR (sg_sweep_csr.cu): sparse couplings, replica-minor state
struct CsrDev {
    const long long* rowptr;  // [n+1]  rows of J^T (= columns of J): whom a flip of `site` touches
    const int* colidx;        // [nnz]
    const float* val;         // [nnz]
    const float* diag;        // [n]    J_ii (0 for the usual models)
    int8_t* spins;            // [n][Rp]
    float* fields;            // [n][Rp]
    int8_t* best_spins;       // [n][Rp]
    int Rp;                   // replicas padded to 32
    int symmetric;            // J == J^T: the energy can be carried incrementally
    const float* h;           // [n] (energy recomputation for asymmetric J)
};
size_t csr_sites_bytes(int n, int n_sweeps);
cudaError_t launch_sweep_csr(const CsrDev& m, const SweepDev& a, bool inject, void* sites_buf,
                             cudaStream_t st);
// fields from the rows of J (rowptr_rows...), then energies; m.spins / m.fields / m.Rp are used
cudaError_t launch_csr_fields(const CsrDev& m, const long long* rowptr_rows, const int* colidx_rows,
                              const float* val_rows, const float* h, int n, int R, float* energy,
                              cudaStream_t st);
cudaError_t launch_to_replica_minor_i8(const int8_t* src, int n, int R, int Rp, int8_t* dst,
                                       cudaStream_t st);
cudaError_t launch_from_replica_minor_i8(const int8_t* src, int n, int R, int Rp, int8_t* dst,
                                         cudaStream_t st);
cudaError_t launch_from_replica_minor_f32(const float* src, int n, int R, int Rp, float* dst,
                                          cudaStream_t st);

// K1-LAT (sg_sweep_lattice.cu): checkerboard multi-spin-coded sweep for 2D +-J lattices
struct LatDev {
    uint32_t* lat;        // [W][n]  bit b of word w = spin of replica 32 w + b (1 = up)
    uint32_t* best_lat;   // [W][n]
    const uint8_t* bond;  // [n] bond code: bit 2d present, bit 2d+1 negative; d = up, down, left, right
    int L;
    int n_bonds;
};
int lattice_sequence_index(int L, int x, int y);
cudaError_t launch_lat_pack(const int8_t* spins, int n, int R, uint32_t* lat, cudaStream_t st);
cudaError_t launch_lat_unpack(const uint32_t* lat, int n, int R, int8_t* spins, cudaStream_t st);
cudaError_t launch_lat_energy(const LatDev& m, const uint32_t* lat, int R, float* energy, cudaStream_t st);
cudaError_t launch_sweep_lattice(const LatDev& m, const SweepDev& a, bool inject, uint64_t* launches,
                                 cudaStream_t st);

// K1-GRP (sg_sweep_groups.cu): block-clique couplings, state in shared memory as group sums
struct GrpDev {
    const int* group_of;      // [n]
    const float* coupling;    // [n_groups]  J_ij = coupling[g] for i != j in group g
    const float* h;           // [n]
    uint32_t* words;          // [W][n] spin bit planes (as in lattice mode)
    uint32_t* best_words;     // [W][n]
    int n_groups;
};
// groups dealt out to P partitions (static per model) for the P x W grid of the partitioned kernel
struct GrpPartDev {
    int P;
    const int* part_of_group;  // [n_groups]
    const int* part_off;       // [P+1] offsets into part_sites
    const int* part_sites;     // [n]   sites ordered by partition
    const int* local_site;     // [n]   index of a site inside its partition
    const int* part_goff;      // [P+1] offsets into part_groups
    const int* part_groups;    // [n_groups] groups ordered by partition
    const int* local_group;    // [n_groups] index of a group inside its partition
    const int* goff;           // [n_groups+1] offsets into gsites
    const int* gsites;         // [n] sites ordered by group
    short* sums;               // [W][n_groups][32] group sums per replica word (scratch)
    int max_sites, max_groups; // largest partition
};
size_t groups_part_scratch_bytes(int n, int n_sweeps, int R, int P);
cudaError_t launch_sweep_groups_part(const GrpDev& m, const GrpPartDev& q, const SweepDev& a, bool inject,
                                     const int* sites, void* scratch, uint64_t* launches,
                                     cudaStream_t st);
size_t groups_smem_bytes(int n, int n_groups);
cudaError_t launch_sweep_groups(const GrpDev& m, const SweepDev& a, bool inject, const int* sites,
                                cudaStream_t st);
cudaError_t launch_groups_energy(const GrpDev& m, const uint32_t* words, int n, int R, float* energy,
                                 cudaStream_t st);
// int32 site table [n_sweeps][n] for the sparse / group kernels (same Philox stream as the others)
cudaError_t launch_sites_table(const SweepDev& a, int* out, cudaStream_t st);

// K1-WOLFF (sg_wolff.cu): the reference's cluster move, one CTA per replica, one launch per sweep
struct WolffDev {
    const float* Jrow;        // [n][n_pad] row-major couplings: Jrow[c][j] = J[c][j]
    const unsigned short* nb_col;  // [n][32] ascending columns with a negative coupling (0xFFFF: none), or null
    const float* nb_val;      // [n][32] their couplings
    int8_t* spins;            // [R][n_pad]
    unsigned long long* accepted;  // [R] += cluster sizes
    const double* temps;      // T(s, r) = temps[s*t_ss + r*t_rs]
    long long t_ss, t_rs;
    const int* sites;         // start site of update k: sites[r*s_rs + s*s_ss + k]
    long long s_rs, s_ss;
    const float* uniforms;    // injected: replica r consumes uniforms[r*u_rs + cursor[r] ...] in order
    long long u_rs, u_len;
    long long* cursor;        // [R] in/out
    int* status;              // set to 1 when a replica's stream ran dry
    unsigned long long seed, sweep_abs;
    int n, n_pad, R, sweep, rep_base;
};
cudaError_t launch_wolff(const WolffDev& a, bool inject, cudaStream_t st);
cudaError_t launch_wolff_neg_count(const float* Jrow, int n, int n_pad, int* max_deg, cudaStream_t st);
cudaError_t launch_wolff_neg_fill(const float* Jrow, int n, int n_pad, unsigned short* nb_col, float* nb_val,
                                  cudaStream_t st);
cudaError_t launch_wolff_record(const float* energy, float* best_energy, const int8_t* spins,
                                int8_t* best_spins, float* trace_row, int n_pad, int R, int track_best,
                                cudaStream_t st);

// K3 (sg_exchange.cu)
struct ExchangeDev {
    int* rep_at;              // [L][K] (global) replica currently at rung k of ladder l
    double* rep_temp;         // [rep_n] temperatures of the local replicas
    const double* ladder;     // [K]
    const float* energy;      // indexed by global replica id ([rep_n] when unsharded)
    unsigned int* attempts;   // [L][K-1]
    unsigned int* accepts;    // [L][K-1]
    const double* uniforms;   // injected [L][K/2] (nearest) / [L][K(K-1)] (all pairs) or null
    unsigned long long seed, round;
    int L, K, parity, inject;
    int method;               // 0 = nearest neighbour (even/odd), 1 = all pairs
    int rep_lo, rep_n;        // local replicas are the global ids [rep_lo, rep_lo + rep_n)
};
cudaError_t launch_exchange(ExchangeDev a, cudaStream_t st);
// ADAPTIVE schedule step on the device (see sg_exchange.cu)
cudaError_t launch_adaptive_temperature(const unsigned long long* accepted, unsigned long long accepted_base,
                                        int n, int sweep, int window, double target, double rate,
                                        double t_final, const double* base_temps, double* state,
                                        double* temps_out, cudaStream_t st);
// operator-form exchange (one ordered pass over adjacent pairs, rows swapped in place)
size_t exchange_chain_header_bytes(int R);
cudaError_t launch_exchange_chain(void* rows, long long row_stride, long long row_bytes, int R,
                                  float* energies, const float* temps, const float* uniforms,
                                  unsigned long long seed, unsigned long long round, void* scratch,
                                  cudaStream_t st);
cudaError_t launch_ladder_init(int* rep_at, double* rep_temp, const double* ladder, int n_global, int K,
                               int rep_lo, int rep_n, cudaStream_t st);
cudaError_t launch_ladder_temps(const int* rep_at, const double* ladder, double* rep_temp, int n_global, int K,
                                int rep_lo, int rep_n, cudaStream_t st);
cudaError_t launch_best_config(const float* energy, int R, const int8_t* rows, int n, int n_pad, float* out_e,
                               int* out_idx, int8_t* out_row, cudaStream_t st);
cudaError_t launch_check_target(const float* energy, int R, int rep_lo, float target, int round, int* hit,
                                cudaStream_t st);

}  // namespace sg
