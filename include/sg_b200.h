/*
 * sg_b200.h -- C ABI of the B200-native Ising annealing engine (libsg_b200.so).
 *
 * Drop-in boundary for ONE hot path of danieleschmidt/spin-glass-anneal-rl: the
 * batched Monte Carlo sweep that IsingModel / GPUAnnealer.anneal() /
 * ParallelTempering.run() drive.  Plain pointers and sizes only; no torch types.
 * The reference is pure Python, so its "FFI" for this path is the set of device
 * entry points it declares but can never launch (annealing/cuda_kernels.py) plus
 * the Python methods that loop around them.  Each entry point below cites the
 * reference interface it replaces (paths relative to /root/reference/spin_glass_rl/).
 *
 * Conventions
 *   - every function returns 0 on success or a negative sg_status; the message
 *     is available from sg_last_error() (thread-local);
 *   - "dev" pointers are CUDA device pointers on the engine's device, "host"
 *     pointers are ordinary host memory; functions taking `on_device` accept
 *     either and copy accordingly;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     all work is enqueued on it and is asynchronous unless a host buffer is
 *     involved, in which case the call returns after the copy completed;
 *   - spins are int8 in {-1,+1}, row-major [R][n]; couplings float32 row-major.
 *   - an engine is not re-entrant; use one engine per host thread / stream
 *     (the reference's callers fan out over distinct models).
 */
#ifndef SG_B200_H
#define SG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_ABI_VERSION 5

typedef enum {
    SG_OK = 0,
    SG_ERR_INVALID = -1, /* bad argument / wrong call order   */
    SG_ERR_CUDA = -2,    /* CUDA runtime error                */
    SG_ERR_NOMEM = -3,   /* allocation failed                 */
    SG_ERR_UNSUPPORTED = -4
} sg_status;

/* update rules: core/spin_dynamics.py:11-16 (UpdateRule) */
#define SG_RULE_METROPOLIS 0 /* _metropolis_update  core/spin_dynamics.py:131-152 */
#define SG_RULE_GLAUBER 1    /* _glauber_update     core/spin_dynamics.py:154-170 */
#define SG_RULE_HEAT_BATH 2  /* _heat_bath_update   core/spin_dynamics.py:172-191 */
#define SG_RULE_WOLFF 3      /* _wolff_update       core/spin_dynamics.py:193-262; a cluster move with
                                its own entry point, sg_sweep_wolff (sg_sweep rejects it)        */

/* where the acceptance randomness comes from */
#define SG_RNG_PHILOX 0   /* in-kernel Philox4x32-10 counter RNG (production)          */
#define SG_RNG_INJECTED 1 /* one caller-supplied float32 uniform per (replica, attempt) */

/* which site each attempt visits (shared by all replicas of a thread block) */
#define SG_SITES_SEQUENTIAL 0 /* 0,1,...,n-1 every sweep                                   */
#define SG_SITES_RANDOM 1     /* Philox draw % n, with replacement (core/spin_dynamics.py:69) */
#define SG_SITES_EXPLICIT 2   /* caller-supplied list (replay of a recorded stream)         */
#define SG_SITES_RANDOM_PER_BLOCK 3 /* like RANDOM, but an independent stream per thread block */
#define SG_SITES_CHECKERBOARD 4     /* lattice models: all (x+y) even sites in row-major order,
                                       then all odd ones -- each site once per sweep          */

/* which sweep kernel runs (sg_sweep_params.kernel) */
#define SG_KERNEL_AUTO 0 /* n <= 224: SMALL; else the tensor-core kernel when the model/launch
                            shape allows it; else SIMT                                          */
#define SG_KERNEL_SIMT 1 /* register-resident fields, sequential fp32 FMAs (sg_sweep.cu)          */
#define SG_KERNEL_TC 2   /* TMEM-resident fields, tcgen05 rank-16 block updates (sg_sweep_tc.cu)  */
#define SG_KERNEL_SMALL 3 /* n <= 224: J in shared memory, one warp per replica, fields in
                             registers (sg_sweep_small.cu)                                      */

typedef struct sg_engine sg_engine;

int sg_abi_version(void);
const char *sg_last_error(void);

/* Engine = one model (J, h) + R replicas resident on one GPU.
 * Replaces the per-call state the reference rebuilds in
 * GPUAnnealer._move_model_to_gpu (annealing/gpu_annealer.py:185-197). */
int sg_create(int device_id, sg_engine **out);
void sg_destroy(sg_engine *e);

/* IsingModel.set_couplings_from_matrix / set_external_fields
 * (core/ising_model.py:106-123).  J is n x n float32 with row stride ldJ; it need
 * not be symmetric: like the reference, the local field of spin i is
 * sum_j J[i][j] s_j + h[i] (diagonal included, core/ising_model.py:176-185). */
int sg_set_model_dense(sg_engine *e, int n, const float *J, int64_t ldJ, const float *h,
                       int on_device, void *stream);

/* Sparse couplings (the models the reference's callers build as torch sparse COO,
 * problems/base.py:107-116; core/ising_model.py:70-80): CSR rows of J, host arrays.  The local
 * field of spin i is sum over row i + h[i], exactly as for dense models.  Replaces the dense model
 * of the engine; the sweep then runs the sparse kernel (sg_sweep_csr.cu): one site order per
 * launch, any n (no 7168 limit), state kept replica-minor.  All other entry points behave the
 * same (sg_tc_selftest and the kernel / coupling_planes fields of sg_sweep_params do not apply). */
int sg_set_model_csr(sg_engine *e, int n, int64_t nnz, const int64_t *rowptr, const int32_t *colidx,
                     const float *val, const float *h, void *stream);

/* 2D +-J lattice (the Edwards-Anderson instances of research/experimental_validation.py:134-180):
 * L x L spins, spin (x, y) = x * L + y, Jx[x*L+y] couples (x, y) with (x+1 mod L, y), Jy[x*L+y]
 * couples (x, y) with (x, y+1 mod L); entries -1, 0 (no bond: open boundaries) or +1; host arrays;
 * h = 0.  Selects the checkerboard multi-spin-coded kernel (sg_sweep_lattice.cu): sg_sweep must
 * use SG_SITES_CHECKERBOARD; local fields are not materialised (sg_get_fields is unsupported),
 * everything else behaves the same.  Periodic bonds need an even L. */
int sg_set_model_lattice2d(sg_engine *e, int L, const int8_t *Jx, const int8_t *Jy, void *stream);
/* position of site (x, y) in the checkerboard attempt sequence of one sweep (replay tests) */
int sg_lattice_sequence_index(int L, int x, int y);

/* Block-clique couplings: J_ij = coupling[g] for every pair i != j of group g, 0 between groups --
 * what the reference's cardinality / one-hot penalties produce (core/constraints.py:126-158; e.g.
 * SimpleScheduler, problems/simple_scheduler.py:67-127: one group per task).  The local field is
 * h_i + c_g (S_g - s_i) with S_g the spin sum of the group, so the sweep keeps only spin bits and
 * integer group sums, in shared memory (sg_sweep_groups.cu); needs 4 n + 64 n_groups bytes <= 227 KB
 * (SG_ERR_UNSUPPORTED otherwise: use sg_set_model_csr).  Host arrays.  Site orders SEQUENTIAL /
 * RANDOM / EXPLICIT; sg_get_fields is unsupported; everything else behaves the same. */
int sg_set_model_groups(sg_engine *e, int n, int n_groups, const int32_t *group_of,
                        const float *coupling, const float *h, void *stream);

/* Several small dense models in one engine (n <= 224 each, same n): J is [n_models][n][n]
 * float32 contiguous, h is [n_models][n].  The replicas allocated afterwards are model-major
 * (n_replicas / n_models each; replica r anneals model r / (n_replicas / n_models)), every
 * launch covers all models at once on the small-model kernel (sg_sweep_small.cu).  This is what
 * BatchProcessor.process_models_batch (annealing/batch_processor.py:231-288, 423-454: one
 * GPUAnnealer.anneal per model on a 4-thread pool) and the RL environment's many short anneals
 * (rl_integration/environment.py:318-336) become.  sg_batch_energies then takes a model-major
 * batch (a multiple of n_models configurations). */
int sg_set_model_dense_batch(sg_engine *e, int n_models, int n, const float *J, const float *h,
                             int on_device, void *stream);

/* Allocate R replicas (spins, local fields, energies, best-so-far, counters). */
int sg_alloc_replicas(sg_engine *e, int n_replicas, void *stream);

/* IsingModel.set_spins / get_spins (core/ising_model.py:191-198), batched. */
int sg_set_spins(sg_engine *e, const int8_t *spins, int on_device, void *stream);
int sg_get_spins(sg_engine *e, int8_t *spins, int on_device, void *stream);

/* Checkpoint / resume of a run in progress (the reference only saves finished
 * results, annealing/result.py:147-188).  Together with sg_set_spins these restore everything a
 * resumed run depends on: best-so-far records, acceptance counters and the ladder state (the
 * per-replica temperatures follow from the rung -> replica map).  Fields and energies are
 * recomputed by sg_init_fields -- call it BEFORE sg_set_best, it resets the best records.  The
 * Philox streams are counter based (seed, absolute sweep index, global replica id), so a run
 * resumed at sweep s continues exactly as the uninterrupted one.  Dense models. */
int sg_set_best(sg_engine *e, const float *best_energy, const int8_t *best_spins, int on_device, void *stream);
int sg_set_accepted(sg_engine *e, const uint64_t *accepted, int on_device, void *stream);
int sg_set_ladder_state(sg_engine *e, const int32_t *replica_at_rung, const uint32_t *attempts,
                        const uint32_t *accepts, int on_device, void *stream);

/* Pipelined upload for callers that stream many start configurations through one engine (the RL
 * environment's repeated anneals, rl_integration/environment.py:318-336; bench.py's end-to-end
 * leg): sg_upload_spins_async copies spins[R][n] from PINNED host memory into one of two
 * engine-owned staging buffers asynchronously on `stream` (which may be a side stream running
 * ahead of the compute stream; the host buffer must stay untouched until that stream reaches
 * this point), sg_set_spins_staged then makes the staged configurations the current spins on the
 * compute stream (the caller orders the two streams with an event).  Dense models. */
int sg_upload_spins_async(sg_engine *e, const int8_t *host_spins, int slot, void *stream);
int sg_set_spins_staged(sg_engine *e, int slot, void *stream);

/* Local-field initialisation + energies for all replicas in one pass:
 *   F = S J^T + h ; E_r = -1/2 sum_i s_ri (F_ri + h_i)
 * replaces R x IsingModel.compute_energy (core/ising_model.py:149-174) and
 * CUDAKernelManager.compute_energy_optimized (annealing/cuda_kernels.py:284-324).
 * Must be called after sg_set_spins and before sg_sweep. */
int sg_init_fields(sg_engine *e, void *stream);

/* Recompute local fields and energies exactly from the current spins WITHOUT touching the
 * best-so-far records: removes the rounding drift the incremental updates accumulate for
 * non-integer couplings (the reference recomputes every local field from scratch,
 * core/ising_model.py:176-185, so it never drifts).  Cheap: one K2 pass. */
int sg_refresh_fields(sg_engine *e, void *stream);

int sg_get_energies(sg_engine *e, float *energies, int on_device, void *stream);
int sg_get_fields(sg_engine *e, float *fields, int on_device, void *stream);
int sg_get_accepted(sg_engine *e, uint64_t *accepted, int on_device, void *stream);

/* Best-so-far per replica (GPUAnnealer best tracking, annealing/gpu_annealer.py:130-131,
 * 151-153: compared once per sweep).  reset sets best = current. */
int sg_reset_best(sg_engine *e, void *stream);
int sg_get_best(sg_engine *e, float *best_energy, int8_t *best_spins, int on_device, void *stream);

/* The one result GPUAnnealer.anneal returns (annealing/gpu_annealer.py:166-183): the lowest
 * best-so-far energy over all replicas, its replica index and that configuration (int8[n]) --
 * an argmin on the device and a 4 + 4 + n byte transfer instead of all R configurations.  Host
 * outputs are written asynchronously (pinned memory: valid once `stream` has been synchronised).
 * Dense models; any output may be NULL. */
int sg_get_best_config(sg_engine *e, float *best_energy, int32_t *replica, int8_t *spins,
                       int on_device, void *stream);

typedef struct {
    uint32_t struct_size;       /* = sizeof(sg_sweep_params)                                  */
    int32_t n_sweeps;           /* sweeps in this launch; one sweep = n attempts per replica   */
    int32_t rule;               /* SG_RULE_*                                                   */
    int32_t rng_mode;           /* SG_RNG_*                                                    */
    int32_t site_mode;          /* SG_SITES_*                                                  */
    int32_t replicas_per_block; /* 0 = auto (largest the register file allows)                */
    /* temperature of replica r in sweep s: temps[s*temps_sweep_stride + r*temps_replica_stride]
     * (dev, float64).  NULL = the engine's per-replica ladder temperatures (sg_set_ladder). */
    const double *temps;
    int64_t temps_sweep_stride;
    int64_t temps_replica_stride;
    uint64_t seed;       /* Philox key                                                        */
    uint64_t sweep_base; /* absolute index of the first sweep (Philox counter; makes results
                            independent of how a run is cut into launches)                   */
    /* SG_SITES_EXPLICIT: site of attempt k of sweep s for block b =
     * sites[b*sites_block_stride + s*sites_sweep_stride + k]   (dev, int32) */
    const int32_t *sites;
    int64_t sites_block_stride;
    int64_t sites_sweep_stride;
    /* SG_RNG_INJECTED: uniforms[(r*n_sweeps + s)*n + k]   (dev, float32) */
    const float *uniforms;
    float *energy_trace; /* optional dev out [n_sweeps][R]: energy after every sweep           */
    int32_t track_best;  /* compare-and-keep best energy/configuration after every sweep       */
    int32_t kernel;      /* SG_KERNEL_*                                                        */
    /* SG_KERNEL_TC: number of bf16 planes the couplings are summed from: 3 = every fp32 coupling
     * exactly, 2 = 16 significant bits, 1 = plain bf16.  0 (default) = as many as THIS model needs
     * to be exact: 3 for arbitrary fp32 couplings, 1 when every coupling is a bf16 value (integer
     * couplings |J| <= 256) -- a third of the tensor-core work, same results. */
    int32_t coupling_planes;
    /* global id of this engine's replica 0: the Philox key of replica r is replica_base + r, so a
     * replica set sharded over several engines / GPUs draws exactly the numbers the same replicas
     * would draw in one engine (0 for a single engine; a multiple of 32 for lattice models) */
    int32_t replica_base;
    /* optional dev out [R][n] (float32, caller-zeroed): every accepted flip adds its energy change
     * 2 s_i f_i to entry (replica, site) -- the third return value of
     * CUDAKernelManager.metropolis_update_optimized (annealing/cuda_kernels.py:228-282, 371-397).
     * Dense models on the sequential-FMA kernel only (kernel = SG_KERNEL_SIMT or replay). */
    float *site_energy_changes;
} sg_sweep_params;

/* The sweep: replaces SpinDynamics.sweep() (core/spin_dynamics.py:73-94) looped over
 * replicas and sweeps, i.e. the body of GPUAnnealer.anneal's loop
 * (annealing/gpu_annealer.py:139-153), ParallelTempering._parallel_sweeps
 * (annealing/parallel_tempering.py:191-203) and the never-launched
 * CUDAKernelManager.metropolis_update_optimized (annealing/cuda_kernels.py:228-282). */
int sg_sweep(sg_engine *e, const sg_sweep_params *p, void *stream);

typedef struct {
    uint32_t struct_size; /* = sizeof(sg_wolff_params)                                           */
    int32_t n_sweeps;     /* one sweep = n cluster updates per replica (SpinDynamics.sweep)       */
    int32_t rng_mode;     /* SG_RNG_*                                                             */
    int32_t site_mode;    /* start sites: SG_SITES_SEQUENTIAL / _RANDOM (one list for all replicas)
                             / _EXPLICIT                                                          */
    /* temperature of replica r in sweep s, addressed like the temps of sg_sweep_params;
     * NULL = the ladder temperatures */
    const double *temps;
    int64_t temps_sweep_stride;
    int64_t temps_replica_stride;
    uint64_t seed;        /* Philox key                                                           */
    uint64_t sweep_base;  /* absolute index of the first sweep (Philox counter)                   */
    /* SG_SITES_EXPLICIT: start site of update k of sweep s for replica r =
     * sites[r*sites_replica_stride + s*sites_sweep_stride + k]   (dev, int32) */
    const int32_t *sites;
    int64_t sites_replica_stride;
    int64_t sites_sweep_stride;
    /* SG_RNG_INJECTED: replica r consumes uniforms[r*uniforms_replica_stride + c], c = cursor[r],
     * cursor[r]+1, ... in the order the reference draws them (one per candidate neighbour: not in
     * the cluster, coupling < 0, same spin -- walked in index order per dequeued site); at most
     * uniforms_per_replica each (SG_ERR_INVALID when a replica needs more; the call then
     * synchronises).  cursor (dev int64 [R], in/out) carries the position across calls. */
    const float *uniforms;
    int64_t uniforms_replica_stride;
    int64_t uniforms_per_replica;
    int64_t *cursor;
    float *energy_trace;  /* optional dev out [n_sweeps][R]: exact energy after every sweep       */
    int32_t track_best;   /* compare-and-keep best energy/configuration after every sweep         */
    int32_t replica_base; /* global id of this engine's replica 0 (Philox key)                    */
} sg_wolff_params;

/* Sweeps of the reference's cluster move, UpdateRule.WOLFF: SpinDynamics.sweep() calling
 * _wolff_update -> _wolff_cluster_dense (core/spin_dynamics.py:73-94, 193-262) n times per sweep,
 * for all replicas (one thread block each, sg_wolff.cu).  Per update: breadth-first growth from
 * the start site over the ROW of every dequeued site, a neighbour joins with probability
 * 1 - exp(2 J / T) if its coupling is negative and its spin equal; the cluster is flipped; the
 * acceptance counter grows by the cluster size (nothing is ever rejected, so the reference's
 * acceptance rate is 1).  External fields do not enter the move.  After every sweep local fields
 * and energies are recomputed exactly (as sg_refresh_fields does).  Dense single models only. */
int sg_sweep_wolff(sg_engine *e, const sg_wolff_params *p, void *stream);

/* Parallel tempering ladder: R replicas = n_ladders x n_rungs; rung 0 is the hottest
 * (ParallelTempering._generate_temperature_ladder, annealing/parallel_tempering.py:146-173).
 * ladder_temps is a host array [n_rungs]. */
int sg_set_ladder(sg_engine *e, int n_rungs, const double *ladder_temps, void *stream);

/* The same for ladders that span several engines / GPUs: the replica set is
 * n_global_replicas = n_ladders x n_rungs replicas with global ids, this engine holds the ids
 * [replica_offset, replica_offset + R).  Every engine keeps the whole rung -> replica map (it
 * evolves identically everywhere because every rank takes the same decisions from the same
 * all-gathered energies and the same counter RNG) and the temperatures of its own replicas.
 * Replaces MultiGPUAnnealer.anneal_replica_exchange (annealing/multi_gpu.py:234-307). */
int sg_set_ladder_sharded(sg_engine *e, int n_rungs, const double *ladder_temps,
                          int n_global_replicas, int replica_offset, void *stream);

typedef struct {
    uint32_t struct_size;
    int32_t parity;          /* first rung of the first pair: 0 or 1 (np.random.randint(0,2)) */
    int32_t rng_mode;        /* SG_RNG_PHILOX or SG_RNG_INJECTED                               */
    int32_t method;          /* SG_EXCHANGE_*                                                  */
    uint64_t seed;
    uint64_t round;          /* Philox counter                                                 */
    /* SG_RNG_INJECTED: dev float64; NEAREST: [n_ladders][n_rungs/2], one per pair;
     * ALL_PAIRS: [n_ladders][n_rungs*(n_rungs-1)], consumed in the order the reference draws
     * them (selection draw of every pair, acceptance draw of the selected ones) */
    const double *uniforms;
    /* sharded ladders: dev float32 [n_global_replicas], energy of every replica by GLOBAL id
     * (the all-gather of sg_get_energies over the ranks); NULL = this engine's own energies */
    const float *energies_all;
} sg_exchange_params;

/* exchange move sets (ParallelTemperingConfig.exchange_method) */
#define SG_EXCHANGE_NEAREST 0   /* even/odd adjacent rungs  annealing/parallel_tempering.py:214-220 */
#define SG_EXCHANGE_ALL_PAIRS 1 /* every pair i<j with probability 0.1, in order (CPU branch of
                                   _all_pairs_exchange, annealing/parallel_tempering.py:228-232);
                                   `parity` is ignored                                            */

/* Replica exchange between adjacent rungs, p = min(1, exp((b_j-b_i)(E_j-E_i))): replaces
 * ParallelTempering._nearest_neighbor_exchange/_attempt_single_exchange
 * (annealing/parallel_tempering.py:214-258) and parallel_tempering_exchange_optimized
 * (annealing/cuda_kernels.py:326-369).  Temperatures move, configurations stay. */
int sg_exchange(sg_engine *e, const sg_exchange_params *p, void *stream);

/* One step of the ADAPTIVE temperature schedule on the device (AdaptiveSchedule.update,
 * annealing/temperature_scheduler.py:206-249; GPUAnnealer feeds it the cumulative acceptance rate
 * of its replica, annealing/gpu_annealer.py:140-147): reads replica `replica`'s acceptance
 * counter, appends rate = (accepted - accepted_base) / (sweep * n) to the ring in `state`
 * (device, window + 1 doubles, zero-initialised by the caller), and writes the temperature of
 * sweep `sweep` to temps_out[sweep]: base_temps[sweep] (device, the geometric base schedule),
 * times (1 - adaptation_rate) / (1 + adaptation_rate) once `window` rates have been seen and
 * their mean is above / below target_acceptance, never below final_temp.  Asynchronous on
 * `stream`: launch it between two sg_sweep calls that take temps_out + sweep as their
 * temperature, and the schedule needs no device -> host read per sweep. */
int sg_adaptive_temperature(sg_engine *e, int replica, uint64_t accepted_base, int sweep, int window,
                            double target_acceptance, double adaptation_rate, double final_temp,
                            const double *base_temps, double *state, double *temps_out, void *stream);

/* The exchange in the shape of the reference's operator,
 * CUDAKernelManager.parallel_tempering_exchange_optimized(spins_arrays, energies, temperatures)
 * (annealing/cuda_kernels.py:326-369; the loop that runs upstream is
 * _parallel_tempering_fallback, :405-436): ONE ordered pass over the pairs (i, i+1),
 * i = 0 .. n_replicas-2, each pair seeing the energies the previous swap left;
 * p = exp((1/T[i+1] - 1/T[i]) * (E[i] - E[i+1])) in float32 as written there, one uniform per
 * pair; an accepted pair swaps rows i and i+1 of `rows` and the two energies IN PLACE
 * (temperatures stay with the index).  `rows` is any device matrix of n_replicas rows of
 * row_bytes bytes, row_stride_bytes apart (float32 or int8 spins); energies, temperatures and
 * the optional uniforms[n_replicas-1] are device float32 (uniforms NULL: Philox keyed on
 * seed / round).  No engine needed.  Synchronises `stream`; *n_accepted (host) = number of
 * accepted pairs. */
int sg_exchange_chain(int device, void *rows, int64_t row_stride_bytes, int64_t row_bytes,
                      int n_replicas, float *energies, const float *temperatures,
                      const float *uniforms, uint64_t seed, uint64_t round, int32_t *n_accepted,
                      void *stream);

/* Early-stop test on the device (time-to-target runs; GPUAnnealer's convergence check reads
 * energies at every record point, annealing/gpu_annealer.py:156-164): if the smallest current
 * (which = 0) or best-so-far (which = 1) energy of this engine's replicas is <= target and
 * hit[0] is still negative, writes hit[1] = that replica's global id, then hit[0] = round.
 * `hit` is an int32[2] the device can write: device memory or pinned host memory (the host can
 * then poll it without synchronising the stream).  Asynchronous. */
int sg_check_target(sg_engine *e, int which, float target, int32_t round, int32_t *hit, void *stream);

/* rung -> replica map [n_global_replicas] (= R unless sharded), per-replica temperature [R], per-pair statistics
 * [n_ladders][n_rungs-1] (ParallelTempering.exchange_attempts/accepts). */
int sg_get_ladder_state(sg_engine *e, int32_t *replica_at_rung, double *replica_temps,
                        uint32_t *attempts, uint32_t *accepts, int on_device, void *stream);

/* Stand-alone batched energies and local fields for arbitrary configurations:
 *   BatchProcessor.process_batch_energies          optimization/high_performance_computing.py:98-165
 *   VectorizedOperations.vectorized_local_fields   optimization/high_performance_computing.py:357-372
 * spins [batch][n] int8; energies [batch] float32 (may be NULL); fields [batch][n] float32
 * (may be NULL). */
int sg_batch_energies(sg_engine *e, int batch, const int8_t *spins, float *energies, float *fields,
                      int on_device, void *stream);

/* Measured streaming bandwidth of this device for a J-sized buffer, in GB/s: one block per
 * SM, every block reads the whole buffer `iters` times with 128-bit loads (the sweep's access
 * pattern; L2-resident when bytes << 126 MB, HBM-bound beyond).  stagger=0: all blocks walk the
 * rows in the same order (as the sweep does); stagger=1: every block starts elsewhere.  This is
 * the roofline denominator bench.py normalises the sweep kernel against. */
int sg_measure_stream_bandwidth(sg_engine *e, int64_t bytes, int iters, int stagger,
                                double *gbps_out);

/* The same measurement for the sweep kernel's actual transport: every block (one per SM)
 * pulls n_rows rows of row_bytes through a depth-stage shared-memory ring with TMA bulk
 * copies (cp.async.bulk + mbarrier), no compute.  GB/s aggregated over all SMs. */
int sg_measure_tma_stream(sg_engine *e, int64_t bytes, int row_bytes, int depth, int n_rows,
                          int stagger, double *gbps_out);

/* Diagnostics for the tensor-core sweep: one rank-16 field update F[r][:] += sum_k
 * deltas[k][r] * J'[sites16[k]][:] on 16 replicas (J' = sum of the first `planes` bf16 planes of
 * the model).  Host buffers: sites16 [16], deltas [16][16] (attempt-major), fields [16][n]. */
int sg_tc_selftest(sg_engine *e, int planes, const int32_t *sites16, const float *deltas,
                   const float *fields_in, float *fields_out);

/* Layout facts the host side needs (padded row length, resident replicas per block...). */
/* CTAs per replica group of the tensor-core sweep for the current model and replica count:
 * 2 = a thread-block cluster pair per 32 replicas, each CTA owning half of the field columns;
 * 1 = one CTA per 16 replicas; 0 = the tensor-core kernel does not take this model. */
int sg_tc_cluster_size(sg_engine *e);
/* With clusters of 4 a GPU keeps 33 of them resident (132 of 148 SMs).  When there are more replica
 * groups than that, a launch of n_sweeps sweeps runs its LAST replicas as cluster pairs on the idle
 * SMs, concurrently (same results: a replica draws the same random numbers and sees the same
 * tensor-core updates in either form).  Returns how many replicas that is for the current model
 * and replica count (0 = none; Philox mode only). */
int sg_tc_side_replicas(sg_engine *e, int n_sweeps, int coupling_planes);

int sg_query(sg_engine *e, int32_t *n, int32_t *n_pad, int32_t *n_replicas,
             int32_t *max_replicas_per_block, int32_t *sm_count);

/* Per-kernel device timing for roofline reporting: while enabled, every sweep-kernel launch (and,
 * for the tensor-core path, every operand-gather launch) is bracketed by CUDA events on its
 * stream.  sg_get_profile synchronises, returns the accumulated milliseconds / launch counts since
 * the last call and resets them. */
int sg_set_profiling(sg_engine *e, int enable);
int sg_get_profile(sg_engine *e, double *sweep_ms, uint64_t *sweep_launches, double *gather_ms,
                   uint64_t *gather_launches);

/* Kernel launch counter (every kernel this library launched on this engine). */
uint64_t sg_launch_count(sg_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* SG_B200_H */
