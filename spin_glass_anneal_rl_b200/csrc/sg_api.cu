// sg_api.cu -- the C ABI of libsg_b200.so (see include/sg_b200.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sg_b200.h"
#include "sg_internal.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* what, cudaError_t ce = cudaSuccess) {
    g_err = what;
    if (ce != cudaSuccess) {
        g_err += ": ";
        g_err += cudaGetErrorName(ce);
        g_err += " (";
        g_err += cudaGetErrorString(ce);
        g_err += ")";
    }
    return code;
}

#define SG_CUDA(call)                                                    \
    do {                                                                 \
        cudaError_t _e = (call);                                         \
        if (_e != cudaSuccess) return fail(SG_ERR_CUDA, #call, _e);      \
    } while (0)

#define SG_REQUIRE(cond, msg)                                   \
    do {                                                        \
        if (!(cond)) return fail(SG_ERR_INVALID, msg);          \
    } while (0)

template <typename T>
int dev_alloc(T** p, size_t count) {
    if (*p) {
        cudaFree(*p);
        *p = nullptr;
    }
    if (count == 0) return SG_OK;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
    if (e != cudaSuccess) {
        *p = nullptr;
        return fail(e == cudaErrorMemoryAllocation ? SG_ERR_NOMEM : SG_ERR_CUDA, "cudaMalloc", e);
    }
    return SG_OK;
}

}  // namespace

struct sg_engine {
    int device = 0;
    int sm_count = 0;
    int n = 0, n_pad = 0, R = 0;
    int n_models = 1;     // stacked small dense models (sg_set_model_dense_batch), K1-SMALL only
    bool stacked = false;
    void* stage = nullptr;   // device staging for host-side model stacks and batch evaluations
    size_t stage_cap = 0;
    float* Jt = nullptr;  // [n][n_pad]
    float* h = nullptr;   // [n_pad]
    void* Jp = nullptr;   // bf16 planes [3][n][n_tc] of Jt for the tensor-core sweep (n <= 4096)
    int n_tc = 0;         // plane row length: n rounded up to 128
    void* dig = nullptr;       // fixed-point int8 digits of Jt, tiled for the K2-TC GEMM
    double* scale = nullptr;   // {2^s, 2^-s} of the fixed-point representation
    unsigned int* info = nullptr;
    void* spin_tiles = nullptr;  // spins of the R replicas in UMMA operand tiles (K2-TC scratch)
    size_t spin_tiles_cap = 0;
    void* tc_sites = nullptr;  // per-launch site tables of the tensor-core sweep
    size_t tc_sites_cap = 0;
    void* tc_stream = nullptr;  // operand stream (gathered J rows in UMMA layout), <= 1 GiB
    size_t tc_stream_cap = 0;
    bool profiling = false;
    sg::KernelTimer timer;
    // 2D +-J lattice mode: multi-spin-coded bit planes, see sg_sweep_lattice.cu
    bool lat = false;
    int l_L = 0, l_bonds = 0;
    // block-clique ("groups") mode shares the bit-plane buffers of lattice mode (lat == true too)
    bool grp = false;
    int g_n = 0;
    int* g_group_of = nullptr;
    float* g_coupling = nullptr;
    // partitions of the groups (P x W grid of the partitioned group kernel)
    int gp_P = 0, gp_max_sites = 0, gp_max_groups = 0;
    int* gp_tables = nullptr;   // part_of_group, part_off, part_sites, local_site, part_goff,
                                // part_groups, local_group, goff, gsites (one allocation)
    short* gp_sums = nullptr;
    size_t gp_sums_cap = 0;
    void* gp_scratch = nullptr;
    size_t gp_scratch_cap = 0;
    uint32_t *l_lat = nullptr, *l_best = nullptr;
    uint8_t* l_bond = nullptr;
    // sparse (CSR) mode: replica-minor state, see sg_sweep_csr.cu
    bool csr = false;
    bool c_symmetric = true;
    long long *c_rowptr_t = nullptr, *c_rowptr_r = nullptr;
    int *c_colidx_t = nullptr, *c_colidx_r = nullptr;
    float *c_val_t = nullptr, *c_val_r = nullptr, *c_diag = nullptr;
    int8_t *c_spins = nullptr, *c_best = nullptr;
    float* c_fields = nullptr;
    int Rp = 0;
    void* c_sites = nullptr;
    size_t c_sites_cap = 0;
    int8_t* spins = nullptr;
    float* fields = nullptr;
    float* energy = nullptr;
    float* best_energy = nullptr;
    int8_t* best_spins = nullptr;
    unsigned long long* accepted = nullptr;
    bool fields_valid = false;
    // ladder
    int K = 0, L = 0;
    int n_global = 0, rep_lo = 0;   // sharded ladders: global replica count, id of local replica 0
    int* rep_at = nullptr;
    double* rep_temp = nullptr;
    double* ladder = nullptr;
    unsigned int* attempts = nullptr;
    unsigned int* accepts = nullptr;
    uint64_t launches = 0;
    long long* dbg = nullptr;
    // staging for asynchronous host transfers (sg_upload_spins_async / sg_get_best_config)
    int8_t* up_stage[2] = {nullptr, nullptr};
    size_t up_cap[2] = {0, 0};
    int* plane_flags = nullptr;          // device [2]: plane 2 / plane 3 of the couplings non-zero
    int planes_needed = 0;               // 0 = not read back yet
    unsigned char* best_out = nullptr;   // device: float energy, int replica, int8 spins[n]
    size_t best_out_cap = 0;
    // Wolff cluster move: row-major copy of the couplings (built on first use), stream-dry flag
    float* Jrow = nullptr;
    int* wolff_status = nullptr;
    unsigned short* wolff_nb_col = nullptr;   // neighbour lists (rows with at most 32 negative couplings)
    float* wolff_nb_val = nullptr;
    int wolff_max_deg = -1;                    // -1: not counted yet for this model
};

namespace {

struct DeviceGuard {
    int prev = 0;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (ok) cudaSetDevice(prev);
    }
};

// copy `bytes` from an engine device buffer to a caller buffer (host or device)
int copy_out(void* dst, const void* src_dev, size_t bytes, int on_device, cudaStream_t st) {
    SG_CUDA(cudaMemcpyAsync(dst, src_dev, bytes,
                            on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!on_device) SG_CUDA(cudaStreamSynchronize(st));
    return SG_OK;
}

void free_replicas(sg_engine* e) {
    cudaFree(e->spins); e->spins = nullptr;
    cudaFree(e->fields); e->fields = nullptr;
    cudaFree(e->energy); e->energy = nullptr;
    cudaFree(e->best_energy); e->best_energy = nullptr;
    cudaFree(e->best_spins); e->best_spins = nullptr;
    cudaFree(e->accepted); e->accepted = nullptr;
    e->R = 0;
    e->fields_valid = false;
}

void free_csr_replicas(sg_engine* e) {
    cudaFree(e->c_spins); e->c_spins = nullptr;
    cudaFree(e->c_best); e->c_best = nullptr;
    cudaFree(e->c_fields); e->c_fields = nullptr;
    e->Rp = 0;
}

void free_csr_model(sg_engine* e) {
    cudaFree(e->c_rowptr_t); e->c_rowptr_t = nullptr;
    cudaFree(e->c_rowptr_r); e->c_rowptr_r = nullptr;
    cudaFree(e->c_colidx_t); e->c_colidx_t = nullptr;
    cudaFree(e->c_colidx_r); e->c_colidx_r = nullptr;
    cudaFree(e->c_val_t); e->c_val_t = nullptr;
    cudaFree(e->c_val_r); e->c_val_r = nullptr;
    cudaFree(e->c_diag); e->c_diag = nullptr;
    e->csr = false;
}

void free_lat_replicas(sg_engine* e) {
    cudaFree(e->l_lat); e->l_lat = nullptr;
    cudaFree(e->l_best); e->l_best = nullptr;
}

void free_lat_model(sg_engine* e) {
    cudaFree(e->l_bond); e->l_bond = nullptr;
    cudaFree(e->g_group_of); e->g_group_of = nullptr;
    cudaFree(e->g_coupling); e->g_coupling = nullptr;
    cudaFree(e->gp_tables); e->gp_tables = nullptr;
    cudaFree(e->gp_sums); e->gp_sums = nullptr; e->gp_sums_cap = 0;
    cudaFree(e->gp_scratch); e->gp_scratch = nullptr; e->gp_scratch_cap = 0;
    e->gp_P = 0;
    e->lat = false;
    e->grp = false;
}

sg::GrpPartDev grp_part_dev(const sg_engine* e) {
    sg::GrpPartDev q{};
    const int n = e->n, G = e->g_n, P = e->gp_P;
    const int* t = e->gp_tables;
    q.P = P;
    q.part_of_group = t;            t += G;
    q.part_off = t;                 t += P + 1;
    q.part_sites = t;               t += n;
    q.local_site = t;               t += n;
    q.part_goff = t;                t += P + 1;
    q.part_groups = t;              t += G;
    q.local_group = t;              t += G;
    q.goff = t;                     t += G + 1;
    q.gsites = t;
    q.sums = e->gp_sums;
    q.max_sites = e->gp_max_sites;
    q.max_groups = e->gp_max_groups;
    return q;
}

sg::GrpDev grp_dev(const sg_engine* e) {
    sg::GrpDev m{};
    m.group_of = e->g_group_of;
    m.coupling = e->g_coupling;
    m.h = e->h;
    m.words = e->l_lat;
    m.best_words = e->l_best;
    m.n_groups = e->g_n;
    return m;
}

sg::LatDev lat_dev(const sg_engine* e) {
    sg::LatDev m{};
    m.lat = e->l_lat;
    m.best_lat = e->l_best;
    m.bond = e->l_bond;
    m.L = e->l_L;
    m.n_bonds = e->l_bonds;
    return m;
}

sg::CsrDev csr_dev(const sg_engine* e) {
    sg::CsrDev m{};
    m.rowptr = e->c_rowptr_t;
    m.colidx = e->c_colidx_t;
    m.val = e->c_val_t;
    m.diag = e->c_diag;
    m.spins = e->c_spins;
    m.fields = e->c_fields;
    m.best_spins = e->c_best;
    m.Rp = e->Rp;
    m.symmetric = e->c_symmetric ? 1 : 0;
    m.h = e->h;
    return m;
}

void free_ladder(sg_engine* e) {
    cudaFree(e->rep_at); e->rep_at = nullptr;
    cudaFree(e->rep_temp); e->rep_temp = nullptr;
    cudaFree(e->ladder); e->ladder = nullptr;
    cudaFree(e->attempts); e->attempts = nullptr;
    cudaFree(e->accepts); e->accepts = nullptr;
    e->K = e->L = 0;
    e->n_global = e->rep_lo = 0;
}

}  // namespace


// ---------------------------------------------------------------- sparse (CSR) mode
namespace {

template <typename T>
int upload(T** dst, const T* src, size_t count, cudaStream_t st) {
    int rc = dev_alloc(dst, count);
    if (rc != SG_OK) return rc;
    if (count) SG_CUDA(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    return SG_OK;
}

int csr_alloc_replicas(sg_engine* e, int n_replicas, cudaStream_t st) {
    free_replicas(e);
    free_csr_replicas(e);
    free_ladder(e);
    const size_t R = (size_t)n_replicas, n = (size_t)e->n;
    const size_t Rp = (R + 31) / 32 * 32;
    int rc;
    if ((rc = dev_alloc(&e->c_spins, n * Rp)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->c_best, n * Rp)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->c_fields, n * Rp)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->energy, R)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->best_energy, R)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->accepted, R)) != SG_OK) return rc;
    SG_CUDA(cudaMemsetAsync(e->c_spins, 1, n * Rp, st));
    SG_CUDA(cudaMemsetAsync(e->c_best, 1, n * Rp, st));
    SG_CUDA(cudaMemsetAsync(e->c_fields, 0, n * Rp * sizeof(float), st));
    SG_CUDA(cudaMemsetAsync(e->accepted, 0, R * sizeof(unsigned long long), st));
    e->R = n_replicas;
    e->Rp = (int)Rp;
    e->fields_valid = false;
    return SG_OK;
}

// caller spins [R][n] (host or device) -> replica-minor buffer dst [n][Rp]
int csr_put_spins(sg_engine* e, const int8_t* spins, int R, int Rp, int8_t* dst, int on_device,
                  cudaStream_t st) {
    const size_t bytes = (size_t)R * e->n;
    const int8_t* src = spins;
    int8_t* tmp = nullptr;
    int rc;
    if (!on_device) {
        if ((rc = dev_alloc(&tmp, bytes)) != SG_OK) return rc;
        SG_CUDA(cudaMemcpyAsync(tmp, spins, bytes, cudaMemcpyHostToDevice, st));
        src = tmp;
    }
    SG_CUDA(sg::launch_to_replica_minor_i8(src, e->n, R, Rp, dst, st));
    e->launches++;
    if (tmp) {
        SG_CUDA(cudaStreamSynchronize(st));
        cudaFree(tmp);
    }
    return SG_OK;
}

int csr_get_i8(sg_engine* e, const int8_t* src_t, int8_t* out, int on_device, cudaStream_t st) {
    const size_t bytes = (size_t)e->R * e->n;
    if (on_device) {
        SG_CUDA(sg::launch_from_replica_minor_i8(src_t, e->n, e->R, e->Rp, out, st));
        e->launches++;
        return SG_OK;
    }
    int8_t* tmp = nullptr;
    int rc;
    if ((rc = dev_alloc(&tmp, bytes)) != SG_OK) return rc;
    cudaError_t ce = sg::launch_from_replica_minor_i8(src_t, e->n, e->R, e->Rp, tmp, st);
    e->launches++;
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, tmp, bytes, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "get spins (csr)", ce);
    return SG_OK;
}

int csr_compute_fields(sg_engine* e, cudaStream_t st) {
    SG_CUDA(sg::launch_csr_fields(csr_dev(e), e->c_rowptr_r, e->c_colidx_r, e->c_val_r, e->h, e->n,
                                  e->R, e->energy, st));
    e->launches += 2;
    return SG_OK;
}

int csr_reset_best(sg_engine* e, cudaStream_t st) {
    SG_CUDA(cudaMemcpyAsync(e->best_energy, e->energy, (size_t)e->R * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
    SG_CUDA(cudaMemcpyAsync(e->c_best, e->c_spins, (size_t)e->n * e->Rp, cudaMemcpyDeviceToDevice, st));
    return SG_OK;
}

int csr_sweep(sg_engine* e, const sg_sweep_params* p, sg::SweepDev a, cudaStream_t st) {
    SG_REQUIRE(p->site_mode != SG_SITES_RANDOM_PER_BLOCK && p->replicas_per_block == 0 &&
                   !(p->site_mode == SG_SITES_EXPLICIT && p->sites_block_stride != 0),
               "sg_sweep (sparse model): one site order per launch, replicas_per_block = 0");
    const size_t need = sg::csr_sites_bytes(e->n, p->n_sweeps);
    if (need > e->c_sites_cap) {
        SG_CUDA(cudaStreamSynchronize(st));
        cudaFree(e->c_sites);
        e->c_sites = nullptr;
        e->c_sites_cap = 0;
        cudaError_t ce = cudaMalloc(&e->c_sites, need);
        if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(site tables)", ce);
        e->c_sites_cap = need;
    }
    a.G = 32;
    if (e->profiling) e->timer.begin(0, st);
    SG_CUDA(sg::launch_sweep_csr(csr_dev(e), a, p->rng_mode == SG_RNG_INJECTED, e->c_sites, st));
    if (e->profiling) e->timer.end(st);
    e->launches += 2;
    return SG_OK;
}

int csr_batch_energies(sg_engine* e, int batch, const int8_t* spins, float* energies, float* fields,
                       int on_device, cudaStream_t st) {
    const int Rp = (batch + 31) / 32 * 32;
    const size_t n = (size_t)e->n;
    int8_t* s_t = nullptr;
    float *f_t = nullptr, *e_dev = nullptr, *f_out = nullptr;
    int rc = SG_OK;
    cudaError_t ce = cudaSuccess;
    do {
        if ((rc = dev_alloc(&s_t, n * Rp)) != SG_OK) break;
        if ((rc = dev_alloc(&f_t, n * Rp)) != SG_OK) break;
        if ((rc = dev_alloc(&e_dev, (size_t)batch)) != SG_OK) break;
        if ((rc = csr_put_spins(e, spins, batch, Rp, s_t, on_device, st)) != SG_OK) break;
        sg::CsrDev m = csr_dev(e);
        m.spins = s_t;
        m.fields = f_t;
        m.Rp = Rp;
        if ((ce = sg::launch_csr_fields(m, e->c_rowptr_r, e->c_colidx_r, e->c_val_r, e->h, e->n, batch,
                                        e_dev, st)))
            break;
        e->launches += 2;
        if (energies &&
            (ce = cudaMemcpyAsync(energies, e_dev, (size_t)batch * sizeof(float),
                                  on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st)))
            break;
        if (fields) {
            float* dstf = fields;
            if (!on_device) {
                if ((rc = dev_alloc(&f_out, (size_t)batch * n)) != SG_OK) break;
                dstf = f_out;
            }
            if ((ce = sg::launch_from_replica_minor_f32(f_t, e->n, batch, Rp, dstf, st))) break;
            e->launches++;
            if (!on_device && (ce = cudaMemcpyAsync(fields, f_out, (size_t)batch * n * sizeof(float),
                                                    cudaMemcpyDeviceToHost, st)))
                break;
        }
        ce = cudaStreamSynchronize(st);
    } while (0);
    cudaFree(s_t);
    cudaFree(f_t);
    cudaFree(e_dev);
    cudaFree(f_out);
    if (rc != SG_OK) return rc;
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "sg_batch_energies (csr)", ce);
    return SG_OK;
}

}  // namespace

extern "C" int sg_set_model_csr(sg_engine* e, int n, int64_t nnz, const int64_t* rowptr,
                                const int32_t* colidx, const float* val, const float* h,
                                void* stream) {
    SG_REQUIRE(e && rowptr && h && (nnz == 0 || (colidx && val)), "sg_set_model_csr: NULL argument");
    SG_REQUIRE(n >= 2 && nnz >= 0 && rowptr[0] == 0 && rowptr[n] == nnz,
               "sg_set_model_csr: need n >= 2 and rowptr[0] = 0, rowptr[n] = nnz");
    for (int64_t k = 0; k < nnz; ++k)
        SG_REQUIRE(colidx[k] >= 0 && colidx[k] < n, "sg_set_model_csr: column index out of range");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // leave dense mode
    free_replicas(e);
    free_csr_replicas(e);
    free_lat_replicas(e);
    free_ladder(e);
    free_csr_model(e);
    free_lat_model(e);
    cudaFree(e->Jt); e->Jt = nullptr;
    cudaFree(e->Jp); e->Jp = nullptr;
    cudaFree(e->dig); e->dig = nullptr;
    // transpose on the host: row i of J^T lists the spins j whose field changes when i flips,
    // with the coupling J[j][i] (the local field of j uses ROW j of J, core/ising_model.py:176-185)
    std::vector<long long> rp_r(rowptr, rowptr + n + 1), rp_t((size_t)n + 1, 0);
    std::vector<int> ci_t((size_t)nnz);
    std::vector<float> v_t((size_t)nnz), diag((size_t)n, 0.0f);
    for (int64_t k = 0; k < nnz; ++k) rp_t[(size_t)colidx[k] + 1]++;
    for (int i = 0; i < n; ++i) rp_t[(size_t)i + 1] += rp_t[(size_t)i];
    {
        std::vector<long long> cur(rp_t.begin(), rp_t.end() - 1);
        for (int j = 0; j < n; ++j)
            for (int64_t k = rowptr[j]; k < rowptr[j + 1]; ++k) {
                const int i = colidx[k];
                const long long pos = cur[(size_t)i]++;
                ci_t[(size_t)pos] = j;
                v_t[(size_t)pos] = val[k];
                if (i == j) diag[(size_t)j] += val[k];
            }
    }
    // symmetric?  (rows of J^T come out sorted by column; compare with the sorted rows of J)
    bool symmetric = true;
    {
        std::vector<std::pair<int, float>> row;
        for (int j = 0; j < n && symmetric; ++j) {
            const int64_t a0 = rowptr[j], a1 = rowptr[j + 1];
            if (rp_t[(size_t)j + 1] - rp_t[(size_t)j] != a1 - a0) { symmetric = false; break; }
            row.clear();
            for (int64_t k = a0; k < a1; ++k) row.emplace_back(colidx[k], val[k]);
            std::sort(row.begin(), row.end());
            for (size_t k = 0; k < row.size(); ++k) {
                const size_t pos = (size_t)rp_t[(size_t)j] + k;
                if (ci_t[pos] != row[k].first || v_t[pos] != row[k].second) { symmetric = false; break; }
            }
        }
    }
    e->c_symmetric = symmetric;
    int rc;
    if ((rc = upload(&e->c_rowptr_r, reinterpret_cast<const long long*>(rp_r.data()), (size_t)n + 1, st)) != SG_OK) return rc;
    if ((rc = upload(&e->c_colidx_r, colidx, (size_t)nnz, st)) != SG_OK) return rc;
    if ((rc = upload(&e->c_val_r, val, (size_t)nnz, st)) != SG_OK) return rc;
    if ((rc = upload(&e->c_rowptr_t, reinterpret_cast<const long long*>(rp_t.data()), (size_t)n + 1, st)) != SG_OK) return rc;
    if ((rc = upload(&e->c_colidx_t, ci_t.data(), (size_t)nnz, st)) != SG_OK) return rc;
    if ((rc = upload(&e->c_val_t, v_t.data(), (size_t)nnz, st)) != SG_OK) return rc;
    if ((rc = upload(&e->c_diag, diag.data(), (size_t)n, st)) != SG_OK) return rc;
    if ((rc = upload(&e->h, h, (size_t)n, st)) != SG_OK) return rc;
    SG_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
    e->n = n;
    e->n_pad = n;
    e->n_tc = 0;
    e->csr = true;
    e->n_models = 1;
    e->stacked = false;
    e->fields_valid = false;
    return SG_OK;
}

// ---------------------------------------------------------------- 2D lattice mode
namespace {

int lat_alloc_replicas(sg_engine* e, int n_replicas, cudaStream_t st) {
    free_replicas(e);
    free_csr_replicas(e);
    free_lat_replicas(e);
    free_ladder(e);
    const size_t R = (size_t)n_replicas, W = (R + 31) / 32, n = (size_t)e->n;
    int rc;
    if ((rc = dev_alloc(&e->l_lat, W * n)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->l_best, W * n)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->energy, R)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->best_energy, R)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->accepted, R)) != SG_OK) return rc;
    SG_CUDA(cudaMemsetAsync(e->l_lat, 0xFF, W * n * sizeof(uint32_t), st));
    SG_CUDA(cudaMemsetAsync(e->l_best, 0xFF, W * n * sizeof(uint32_t), st));
    SG_CUDA(cudaMemsetAsync(e->accepted, 0, R * sizeof(unsigned long long), st));
    e->R = n_replicas;
    e->fields_valid = false;
    return SG_OK;
}

int lat_put_spins(sg_engine* e, const int8_t* spins, int R, uint32_t* dst, int on_device,
                  cudaStream_t st) {
    const size_t bytes = (size_t)R * e->n;
    const int8_t* src = spins;
    int8_t* tmp = nullptr;
    int rc;
    if (!on_device) {
        if ((rc = dev_alloc(&tmp, bytes)) != SG_OK) return rc;
        SG_CUDA(cudaMemcpyAsync(tmp, spins, bytes, cudaMemcpyHostToDevice, st));
        src = tmp;
    }
    SG_CUDA(sg::launch_lat_pack(src, e->n, R, dst, st));
    e->launches++;
    if (tmp) {
        SG_CUDA(cudaStreamSynchronize(st));
        cudaFree(tmp);
    }
    return SG_OK;
}

int lat_get_spins(sg_engine* e, const uint32_t* src, int8_t* out, int on_device, cudaStream_t st) {
    const size_t bytes = (size_t)e->R * e->n;
    if (on_device) {
        SG_CUDA(sg::launch_lat_unpack(src, e->n, e->R, out, st));
        e->launches++;
        return SG_OK;
    }
    int8_t* tmp = nullptr;
    int rc;
    if ((rc = dev_alloc(&tmp, bytes)) != SG_OK) return rc;
    cudaError_t ce = sg::launch_lat_unpack(src, e->n, e->R, tmp, st);
    e->launches++;
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, tmp, bytes, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "get spins (lattice)", ce);
    return SG_OK;
}

int lat_compute_energy(sg_engine* e, cudaStream_t st) {
    if (e->grp) {
        SG_CUDA(sg::launch_groups_energy(grp_dev(e), e->l_lat, e->n, e->R, e->energy, st));
        e->launches++;
        return SG_OK;
    }
    SG_CUDA(sg::launch_lat_energy(lat_dev(e), e->l_lat, e->R, e->energy, st));
    e->launches++;
    return SG_OK;
}

int lat_reset_best(sg_engine* e, cudaStream_t st) {
    const size_t W = ((size_t)e->R + 31) / 32;
    SG_CUDA(cudaMemcpyAsync(e->best_energy, e->energy, (size_t)e->R * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
    SG_CUDA(cudaMemcpyAsync(e->l_best, e->l_lat, W * e->n * sizeof(uint32_t),
                            cudaMemcpyDeviceToDevice, st));
    return SG_OK;
}

int lat_sweep(sg_engine* e, const sg_sweep_params* p, sg::SweepDev a, cudaStream_t st) {
    if (e->grp) {
        SG_REQUIRE(p->site_mode <= SG_SITES_EXPLICIT && p->replicas_per_block == 0 &&
                       !(p->site_mode == SG_SITES_EXPLICIT && p->sites_block_stride != 0),
                   "sg_sweep (group model): one site order per launch, replicas_per_block = 0");
        const size_t need = sg::csr_sites_bytes(e->n, p->n_sweeps);
        if (need > e->c_sites_cap) {
            SG_CUDA(cudaStreamSynchronize(st));
            cudaFree(e->c_sites);
            e->c_sites = nullptr;
            e->c_sites_cap = 0;
            cudaError_t ce = cudaMalloc(&e->c_sites, need);
            if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(site tables)", ce);
            e->c_sites_cap = need;
        }
        SG_CUDA(sg::launch_sites_table(a, static_cast<int*>(e->c_sites), st));
        e->launches += 1;
        // the groups are spread over P partitions (P x W one-warp CTAs, several resident per SM)
        // instead of one warp per 32 replicas holding the whole model (one warp per SM):
        // 14x faster at 1024 replicas, 6x at 4096 (tools/grp_perf.py); SG_GRP_PART=0 forces the
        // one-warp-per-word kernel
        const int W = (e->R + 31) / 32;
        bool part = e->gp_P >= 2;
        if (const char* env = getenv("SG_GRP_PART")) part = e->gp_P >= 2 && atoi(env) != 0;
        if (part) {
            const size_t need_sums = (size_t)W * e->g_n * 32 * sizeof(short);
            const size_t need_scr = sg::groups_part_scratch_bytes(e->n, p->n_sweeps, e->R, e->gp_P);
            if (need_sums > e->gp_sums_cap || need_scr > e->gp_scratch_cap) {
                SG_CUDA(cudaStreamSynchronize(st));
                if (need_sums > e->gp_sums_cap) {
                    cudaFree(e->gp_sums); e->gp_sums = nullptr; e->gp_sums_cap = 0;
                    cudaError_t ce = cudaMalloc(&e->gp_sums, need_sums);
                    if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(group sums)", ce);
                    e->gp_sums_cap = need_sums;
                }
                if (need_scr > e->gp_scratch_cap) {
                    cudaFree(e->gp_scratch); e->gp_scratch = nullptr; e->gp_scratch_cap = 0;
                    cudaError_t ce = cudaMalloc(&e->gp_scratch, need_scr);
                    if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(attempt lists)", ce);
                    e->gp_scratch_cap = need_scr;
                }
            }
            if (e->profiling) e->timer.begin(0, st);
            SG_CUDA(sg::launch_sweep_groups_part(grp_dev(e), grp_part_dev(e), a,
                                                 p->rng_mode == SG_RNG_INJECTED,
                                                 static_cast<const int*>(e->c_sites), e->gp_scratch,
                                                 &e->launches, st));
            if (e->profiling) e->timer.end(st);
            return SG_OK;
        }
        if (e->profiling) e->timer.begin(0, st);
        SG_CUDA(sg::launch_sweep_groups(grp_dev(e), a, p->rng_mode == SG_RNG_INJECTED,
                                        static_cast<const int*>(e->c_sites), st));
        if (e->profiling) e->timer.end(st);
        e->launches += 1;
        return SG_OK;
    }
    SG_REQUIRE(p->site_mode == SG_SITES_CHECKERBOARD,
               "sg_sweep (lattice model): site_mode must be SG_SITES_CHECKERBOARD");
    if (e->profiling) e->timer.begin(0, st);
    SG_CUDA(sg::launch_sweep_lattice(lat_dev(e), a, p->rng_mode == SG_RNG_INJECTED, &e->launches, st));
    if (e->profiling) e->timer.end(st);
    return SG_OK;
}

int lat_batch_energies(sg_engine* e, int batch, const int8_t* spins, float* energies, float* fields,
                       int on_device, cudaStream_t st) {
    SG_REQUIRE(!fields, "sg_batch_energies (lattice model): local fields are not materialised");
    const size_t W = ((size_t)batch + 31) / 32;
    uint32_t* tmp_lat = nullptr;
    float* e_dev = nullptr;
    int rc = SG_OK;
    cudaError_t ce = cudaSuccess;
    do {
        if ((rc = dev_alloc(&tmp_lat, W * e->n)) != SG_OK) break;
        if ((rc = dev_alloc(&e_dev, (size_t)batch)) != SG_OK) break;
        if ((rc = lat_put_spins(e, spins, batch, tmp_lat, on_device, st)) != SG_OK) break;
        if (e->grp) {
            if ((ce = sg::launch_groups_energy(grp_dev(e), tmp_lat, e->n, batch, e_dev, st))) break;
        } else if ((ce = sg::launch_lat_energy(lat_dev(e), tmp_lat, batch, e_dev, st))) {
            break;
        }
        e->launches++;
        if (energies &&
            (ce = cudaMemcpyAsync(energies, e_dev, (size_t)batch * sizeof(float),
                                  on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st)))
            break;
        ce = cudaStreamSynchronize(st);
    } while (0);
    cudaFree(tmp_lat);
    cudaFree(e_dev);
    if (rc != SG_OK) return rc;
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "sg_batch_energies (lattice)", ce);
    return SG_OK;
}

}  // namespace

extern "C" int sg_set_model_lattice2d(sg_engine* e, int L, const int8_t* Jx, const int8_t* Jy,
                                      void* stream) {
    SG_REQUIRE(e && Jx && Jy, "sg_set_model_lattice2d: NULL argument");
    SG_REQUIRE(L >= 2 && L <= 4096, "sg_set_model_lattice2d: need 2 <= L <= 4096");
    const int n = L * L;
    std::vector<uint8_t> bond((size_t)n, 0);
    int n_bonds = 0;
    bool wraps = false;
    auto put = [&](int site, int d, int v) {
        if (v != 0) bond[(size_t)site] |= (uint8_t)((1u << (2 * d)) | ((v < 0 ? 1u : 0u) << (2 * d + 1)));
    };
    for (int x = 0; x < L; ++x)
        for (int y = 0; y < L; ++y) {
            const int s = x * L + y;
            const int jx = Jx[s], jy = Jy[s];
            SG_REQUIRE(jx >= -1 && jx <= 1 && jy >= -1 && jy <= 1,
                       "sg_set_model_lattice2d: couplings must be -1, 0 or +1");
            if (jx != 0) {   // (x, y) -- (x + 1 mod L, y)
                ++n_bonds;
                wraps |= (x == L - 1);
                put(s, 1, jx);
                put(((x + 1) % L) * L + y, 0, jx);
            }
            if (jy != 0) {   // (x, y) -- (x, y + 1 mod L)
                ++n_bonds;
                wraps |= (y == L - 1);
                put(s, 3, jy);
                put(x * L + (y + 1) % L, 2, jy);
            }
        }
    SG_REQUIRE(!(wraps && (L & 1)), "sg_set_model_lattice2d: periodic bonds need an even L (checkerboard)");
    SG_REQUIRE(!(wraps && L == 2), "sg_set_model_lattice2d: periodic bonds need L > 2");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    free_replicas(e);
    free_csr_replicas(e);
    free_lat_replicas(e);
    free_ladder(e);
    free_csr_model(e);
    free_lat_model(e);
    cudaFree(e->Jt); e->Jt = nullptr;
    cudaFree(e->Jp); e->Jp = nullptr;
    cudaFree(e->dig); e->dig = nullptr;
    int rc;
    if ((rc = upload(&e->l_bond, bond.data(), (size_t)n, st)) != SG_OK) return rc;
    SG_CUDA(cudaStreamSynchronize(st));
    e->n = n;
    e->n_pad = n;
    e->n_tc = 0;
    e->l_L = L;
    e->l_bonds = n_bonds;
    e->lat = true;
    e->n_models = 1;
    e->stacked = false;
    e->fields_valid = false;
    return SG_OK;
}

extern "C" int sg_set_model_groups(sg_engine* e, int n, int n_groups, const int32_t* group_of,
                                   const float* coupling, const float* h, void* stream) {
    SG_REQUIRE(e && group_of && coupling && h, "sg_set_model_groups: NULL argument");
    SG_REQUIRE(n >= 2 && n_groups >= 1 && n_groups <= n, "sg_set_model_groups: bad sizes");
    for (int i = 0; i < n; ++i)
        SG_REQUIRE(group_of[i] >= 0 && group_of[i] < n_groups, "sg_set_model_groups: group id out of range");
    if (sg::groups_smem_bytes(n, n_groups) > 227 * 1024)
        return fail(SG_ERR_UNSUPPORTED,
                    "sg_set_model_groups: 4 n + 64 n_groups bytes must fit in 227 KB of shared memory "
                    "(use sg_set_model_csr)");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    free_replicas(e);
    free_csr_replicas(e);
    free_lat_replicas(e);
    free_ladder(e);
    free_csr_model(e);
    free_lat_model(e);
    cudaFree(e->Jt); e->Jt = nullptr;
    cudaFree(e->Jp); e->Jp = nullptr;
    cudaFree(e->dig); e->dig = nullptr;
    int rc;
    if ((rc = upload(&e->g_group_of, reinterpret_cast<const int*>(group_of), (size_t)n, st)) != SG_OK) return rc;
    if ((rc = upload(&e->g_coupling, coupling, (size_t)n_groups, st)) != SG_OK) return rc;
    if ((rc = upload(&e->h, h, (size_t)n, st)) != SG_OK) return rc;
    {
        // partitions: groups dealt out largest first to the least loaded of P <= 32 partitions
        const int G = n_groups, P = G < 32 ? G : 32;
        std::vector<int> gsize(G, 0), order(G), part_of(G), load(P, 0), cnt_g(P, 0);
        for (int i = 0; i < n; ++i) gsize[group_of[i]]++;
        for (int g = 0; g < G; ++g) order[g] = g;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return gsize[x] > gsize[y]; });
        for (int g : order) {
            int best = 0;
            for (int q = 1; q < P; ++q)
                if (load[q] < load[best]) best = q;
            part_of[g] = best;
            load[best] += gsize[g];
            cnt_g[best]++;
        }
        std::vector<int> tab((size_t)3 * G + 2 * (P + 1) + 3 * (size_t)n + G + 1);
        int* t = tab.data();
        int* part_of_group = t;  t += G;
        int* part_off = t;       t += P + 1;
        int* part_sites = t;     t += n;
        int* local_site = t;     t += n;
        int* part_goff = t;      t += P + 1;
        int* part_groups = t;    t += G;
        int* local_group = t;    t += G;
        int* goff = t;           t += G + 1;
        int* gsites = t;
        part_off[0] = part_goff[0] = 0;
        for (int q = 0; q < P; ++q) {
            part_off[q + 1] = part_off[q] + load[q];
            part_goff[q + 1] = part_goff[q] + cnt_g[q];
        }
        std::vector<int> fill_s(part_off, part_off + P), fill_g(part_goff, part_goff + P);
        for (int g = 0; g < G; ++g) {
            part_of_group[g] = part_of[g];
            local_group[g] = fill_g[part_of[g]] - part_goff[part_of[g]];
            part_groups[fill_g[part_of[g]]++] = g;
        }
        for (int i = 0; i < n; ++i) {
            const int q = part_of[group_of[i]];
            local_site[i] = fill_s[q] - part_off[q];
            part_sites[fill_s[q]++] = i;
        }
        goff[0] = 0;
        for (int g = 0; g < G; ++g) goff[g + 1] = goff[g] + gsize[g];
        std::vector<int> fill(goff, goff + G);
        for (int i = 0; i < n; ++i) gsites[fill[group_of[i]]++] = i;
        e->gp_P = P;
        e->gp_max_sites = *std::max_element(load.begin(), load.end());
        e->gp_max_groups = *std::max_element(cnt_g.begin(), cnt_g.end());
        if ((rc = upload(&e->gp_tables, tab.data(), tab.size(), st)) != SG_OK) return rc;
        SG_CUDA(cudaStreamSynchronize(st));   // tab goes out of scope
    }
    SG_CUDA(cudaStreamSynchronize(st));
    e->n = n;
    e->n_pad = n;
    e->n_tc = 0;
    e->g_n = n_groups;
    e->n_models = 1;
    e->stacked = false;
    e->lat = true;
    e->grp = true;
    e->fields_valid = false;
    return SG_OK;
}

extern "C" int sg_lattice_sequence_index(int L, int x, int y) {
    return sg::lattice_sequence_index(L, x, y);
}

extern "C" {

int sg_abi_version(void) { return SG_ABI_VERSION; }

const char* sg_last_error(void) { return g_err.c_str(); }

int sg_create(int device_id, sg_engine** out) {
    if (!out) return fail(SG_ERR_INVALID, "sg_create: out is NULL");
    *out = nullptr;
    int count = 0;
    SG_CUDA(cudaGetDeviceCount(&count));
    if (device_id < 0 || device_id >= count) return fail(SG_ERR_INVALID, "sg_create: bad device id");
    cudaDeviceProp prop;
    SG_CUDA(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major < 10)
        return fail(SG_ERR_UNSUPPORTED, "sg_create: this library is built for sm_100a (B200) only");
    sg_engine* e = new (std::nothrow) sg_engine();
    if (!e) return fail(SG_ERR_NOMEM, "sg_create: host allocation failed");
    e->device = device_id;
    e->sm_count = prop.multiProcessorCount;
    *out = e;
    return SG_OK;
}

void sg_destroy(sg_engine* e) {
    if (!e) return;
    DeviceGuard g(e->device);
    free_replicas(e);
    free_csr_replicas(e);
    free_csr_model(e);
    free_lat_replicas(e);
    free_lat_model(e);
    cudaFree(e->c_sites);
    free_ladder(e);
    cudaFree(e->Jt);
    cudaFree(e->h);
    cudaFree(e->Jp);
    cudaFree(e->tc_sites);
    cudaFree(e->tc_stream);
    cudaFree(e->stage);
    cudaFree(e->up_stage[0]);
    cudaFree(e->up_stage[1]);
    cudaFree(e->best_out);
    cudaFree(e->Jrow);
    cudaFree(e->wolff_status);
    cudaFree(e->wolff_nb_col);
    cudaFree(e->wolff_nb_val);
    cudaFree(e->plane_flags);
    cudaFree(e->dig);
    cudaFree(e->scale);
    cudaFree(e->info);
    cudaFree(e->spin_tiles);
    delete e;
}

// Row length of the bf16 planes / fixed-point digits: n rounded up to 128, or to 1024 when that
// costs at most 40 % more columns -- cluster pairs need a multiple of 1024 (each CTA's half of
// the columns in whole 4-tile chunks) and are 1.6x faster per column.  It may exceed the padded
// row n_pad of the replica arrays; the kernels guard every access beyond it.
static int tc_row_length(int n) {
    const int n128 = (n + 127) / 128 * 128;
    const int n1024 = (n + 1023) / 1024 * 1024;
    return (n1024 <= 4096 && 5 * n1024 <= 7 * n128) ? n1024 : n128;
}

int sg_set_model_dense(sg_engine* e, int n, const float* J, int64_t ldJ, const float* h,
                       int on_device, void* stream) {
    SG_REQUIRE(e && J && h, "sg_set_model_dense: NULL argument");
    SG_REQUIRE(n >= 2 && ldJ >= n, "sg_set_model_dense: need n >= 2 and ldJ >= n");
    // padded row length: the smallest size the sweep kernel is instantiated for
    int n_pad = sg::kColQuantum;
    while (n_pad < n) n_pad += sg::kColQuantum;
    if (sg::sweep_max_replicas_per_block(n_pad) == 0)
        return fail(SG_ERR_UNSUPPORTED, "sg_set_model_dense: dense models support n <= 7168");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n != e->n || e->csr || e->lat) {
        free_replicas(e);
        free_csr_replicas(e);
        free_lat_replicas(e);
        free_ladder(e);
    }
    free_csr_model(e);
    free_lat_model(e);
    if (e->stacked) {
        free_replicas(e);
        free_ladder(e);
    }
    e->n_models = 1;
    e->stacked = false;
    e->n = n;
    e->n_pad = n_pad;
    cudaFree(e->Jrow);
    e->Jrow = nullptr;
    cudaFree(e->wolff_nb_col);
    e->wolff_nb_col = nullptr;
    cudaFree(e->wolff_nb_val);
    e->wolff_nb_val = nullptr;
    e->wolff_max_deg = -1;
    int rc;
    if ((rc = dev_alloc(&e->Jt, (size_t)n * n_pad)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->h, (size_t)n_pad)) != SG_OK) return rc;
    const float* Jdev = J;
    float* tmp = nullptr;
    if (!on_device) {
        if ((rc = dev_alloc(&tmp, (size_t)n * ldJ)) != SG_OK) return rc;
        SG_CUDA(cudaMemcpyAsync(tmp, J, (size_t)n * ldJ * sizeof(float), cudaMemcpyHostToDevice, st));
        Jdev = tmp;
    }
    SG_CUDA(sg::launch_pad_transpose(Jdev, ldJ, n, e->Jt, n_pad, st));
    e->launches++;
    SG_CUDA(cudaMemsetAsync(e->h, 0, (size_t)n_pad * sizeof(float), st));
    SG_CUDA(cudaMemcpyAsync(e->h, h, (size_t)n * sizeof(float),
                            on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    // fixed-point digits for the exact tensor-core field initialisation (K2-TC)
    {
        const int n_tc = tc_row_length(n);
        cudaFree(e->dig);
        e->dig = nullptr;
        if (!e->scale) SG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->scale), 2 * sizeof(double)));
        if (!e->info) SG_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->info), 2 * sizeof(unsigned int)));
        cudaError_t ce = cudaMalloc(&e->dig, sg::fields_tc_digits_bytes(n, n_tc));
        if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(digits)", ce);
        SG_CUDA(sg::launch_fields_tc_prepare(e->Jt, n, n_pad, n_tc, e->info, e->scale, e->dig, st));
        e->launches += 3;
    }
    // bf16 planes for the tensor-core sweep (16 replicas x n_tc fp32 fields fill the TMEM)
    cudaFree(e->Jp);
    e->Jp = nullptr;
    e->n_tc = tc_row_length(n);
    if (n <= 4096) {
        const int n_tc = tc_row_length(n);
        void* jp = nullptr;
        cudaError_t ce = cudaMalloc(&jp, (size_t)3 * n * n_tc * 2);
        if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(planes)", ce);
        e->Jp = jp;
        if (!e->plane_flags) {
            int rc2 = dev_alloc(&e->plane_flags, (size_t)2);
            if (rc2 != SG_OK) return rc2;
        }
        SG_CUDA(cudaMemsetAsync(e->plane_flags, 0, 2 * sizeof(int), st));
        SG_CUDA(sg::launch_split_planes(e->Jt, n, n_pad, e->Jp, n_tc, e->plane_flags, st));
        e->planes_needed = 0;   // read back on the first sweep that leaves the choice to the engine
        e->launches++;
    }
    if (tmp) {
        SG_CUDA(cudaStreamSynchronize(st));
        cudaFree(tmp);
    }
    e->fields_valid = false;
    return SG_OK;
}

int sg_set_model_dense_batch(sg_engine* e, int n_models, int n, const float* J, const float* h,
                             int on_device, void* stream) {
    SG_REQUIRE(e && J && h, "sg_set_model_dense_batch: NULL argument");
    SG_REQUIRE(n_models >= 1 && n >= 2, "sg_set_model_dense_batch: need n_models >= 1 and n >= 2");
    if (!sg::sweep_small_supported(n))
        return fail(SG_ERR_UNSUPPORTED, "sg_set_model_dense_batch: stacked models need n <= 224");
    int n_pad = sg::kColQuantum;
    while (n_pad < n) n_pad += sg::kColQuantum;
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t M = (size_t)n_models;
    int rc;
    // a stack of the same shape as the current one (the RL loop: new couplings, same sizes)
    // reuses every buffer, replicas included; cudaMalloc / cudaFree cost milliseconds each
    const bool same = e->stacked && e->n_models == n_models && e->n == n && e->Jt && e->h && !e->csr && !e->lat;
    if (!same) {
        free_replicas(e);
        free_csr_replicas(e);
        free_lat_replicas(e);
        free_ladder(e);
        free_csr_model(e);
        free_lat_model(e);
        cudaFree(e->dig); e->dig = nullptr;
        cudaFree(e->Jp); e->Jp = nullptr;
        e->n = n;
        e->n_pad = n_pad;
        e->n_tc = 0;
        e->n_models = n_models;
        e->stacked = true;
        if ((rc = dev_alloc(&e->Jt, M * n * n_pad)) != SG_OK) return rc;
        if ((rc = dev_alloc(&e->h, M * n_pad)) != SG_OK) return rc;
    }
    const float *Jdev = J, *hdev = h;
    if (!on_device) {
        const size_t need = (M * n * n + M * n) * sizeof(float);
        if (need > e->stage_cap) {
            SG_CUDA(cudaStreamSynchronize(st));
            cudaFree(e->stage); e->stage = nullptr; e->stage_cap = 0;
            cudaError_t ce = cudaMalloc(&e->stage, need);
            if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(model staging)", ce);
            e->stage_cap = need;
        }
        float* tmp = static_cast<float*>(e->stage);
        SG_CUDA(cudaMemcpyAsync(tmp, J, M * n * n * sizeof(float), cudaMemcpyHostToDevice, st));
        SG_CUDA(cudaMemcpyAsync(tmp + M * n * n, h, M * n * sizeof(float), cudaMemcpyHostToDevice, st));
        Jdev = tmp;
        hdev = tmp + M * n * n;
    }
    SG_CUDA(sg::launch_stack_models(Jdev, hdev, n_models, n, n_pad, e->Jt, e->h, st));
    e->launches += 1;
    if (!on_device) SG_CUDA(cudaStreamSynchronize(st));   // the host arrays may be reused by the caller
    e->fields_valid = false;
    return SG_OK;
}

int sg_alloc_replicas(sg_engine* e, int n_replicas, void* stream) {
    SG_REQUIRE(e && e->n > 0, "sg_alloc_replicas: set the model first");
    SG_REQUIRE(n_replicas >= 1, "sg_alloc_replicas: need n_replicas >= 1");
    SG_REQUIRE(!e->stacked || e->csr || e->lat || n_replicas % e->n_models == 0,
               "sg_alloc_replicas: stacked models need the same number of replicas each");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (e->csr) return csr_alloc_replicas(e, n_replicas, st);
    if (e->lat) return lat_alloc_replicas(e, n_replicas, st);
    const size_t R = (size_t)n_replicas, np = (size_t)e->n_pad;
    int rc;
    if (e->R == n_replicas && e->spins && e->fields && e->accepted) {
        // same shape as before (repeated anneals of same-sized models): keep the buffers
        free_ladder(e);
        SG_CUDA(cudaMemsetAsync(e->fields, 0, R * np * sizeof(float), st));
        SG_CUDA(cudaMemsetAsync(e->spins, 1, R * np, st));
        SG_CUDA(cudaMemsetAsync(e->best_spins, 1, R * np, st));
        SG_CUDA(cudaMemsetAsync(e->accepted, 0, R * sizeof(unsigned long long), st));
        e->fields_valid = false;
        return SG_OK;
    }
    free_replicas(e);
    free_ladder(e);
    if ((rc = dev_alloc(&e->spins, R * np)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->fields, R * np)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->energy, R)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->best_energy, R)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->best_spins, R * np)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->accepted, R)) != SG_OK) return rc;
    SG_CUDA(cudaMemsetAsync(e->fields, 0, R * np * sizeof(float), st));
    SG_CUDA(cudaMemsetAsync(e->spins, 1, R * np, st));
    SG_CUDA(cudaMemsetAsync(e->best_spins, 1, R * np, st));
    SG_CUDA(cudaMemsetAsync(e->accepted, 0, R * sizeof(unsigned long long), st));
    e->R = n_replicas;
    e->fields_valid = false;
    return SG_OK;
}

int sg_set_spins(sg_engine* e, const int8_t* spins, int on_device, void* stream) {
    SG_REQUIRE(e && spins && e->R > 0, "sg_set_spins: allocate replicas first");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (e->csr) {
        e->fields_valid = false;
        return csr_put_spins(e, spins, e->R, e->Rp, e->c_spins, on_device, st);
    }
    if (e->lat) {
        e->fields_valid = false;
        return lat_put_spins(e, spins, e->R, e->l_lat, on_device, st);
    }
    const size_t bytes = (size_t)e->R * e->n;
    const int8_t* src = spins;
    int8_t* tmp = nullptr;
    int rc;
    if (!on_device) {
        if ((rc = dev_alloc(&tmp, bytes)) != SG_OK) return rc;
        SG_CUDA(cudaMemcpyAsync(tmp, spins, bytes, cudaMemcpyHostToDevice, st));
        src = tmp;
    }
    SG_CUDA(sg::launch_pad_spins(src, e->n, e->spins, e->n_pad, e->R, st));
    e->launches++;
    if (tmp) {
        SG_CUDA(cudaStreamSynchronize(st));
        cudaFree(tmp);
    }
    e->fields_valid = false;
    return SG_OK;
}

static int get_unpadded_i8(sg_engine* e, const int8_t* src_pad, int8_t* out, int on_device,
                           cudaStream_t st) {
    const size_t bytes = (size_t)e->R * e->n;
    int rc;
    if (on_device) {
        SG_CUDA(sg::launch_unpad_spins(src_pad, e->n_pad, out, e->n, e->R, st));
        e->launches++;
        return SG_OK;
    }
    int8_t* tmp = nullptr;
    if ((rc = dev_alloc(&tmp, bytes)) != SG_OK) return rc;
    cudaError_t ce = sg::launch_unpad_spins(src_pad, e->n_pad, tmp, e->n, e->R, st);
    e->launches++;
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, tmp, bytes, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "get spins", ce);
    return SG_OK;
}

int sg_get_spins(sg_engine* e, int8_t* spins, int on_device, void* stream) {
    SG_REQUIRE(e && spins && e->R > 0, "sg_get_spins: allocate replicas first");
    DeviceGuard g(e->device);
    if (e->csr) return csr_get_i8(e, e->c_spins, spins, on_device, static_cast<cudaStream_t>(stream));
    if (e->lat) return lat_get_spins(e, e->l_lat, spins, on_device, static_cast<cudaStream_t>(stream));
    return get_unpadded_i8(e, e->spins, spins, on_device, static_cast<cudaStream_t>(stream));
}

int sg_reset_best(sg_engine* e, void* stream) {
    SG_REQUIRE(e && e->R > 0 && e->fields_valid, "sg_reset_best: call sg_init_fields first");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (e->csr) return csr_reset_best(e, st);
    if (e->lat) return lat_reset_best(e, st);
    SG_CUDA(cudaMemcpyAsync(e->best_energy, e->energy, (size_t)e->R * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
    SG_CUDA(cudaMemcpyAsync(e->best_spins, e->spins, (size_t)e->R * e->n_pad,
                            cudaMemcpyDeviceToDevice, st));
    return SG_OK;
}

static int compute_fields(sg_engine* e, cudaStream_t st) {
    if (e->csr) return csr_compute_fields(e, st);
    if (e->lat) return lat_compute_energy(e, st);
    if (e->stacked) {
        SG_CUDA(sg::launch_fields_small(e->Jt, e->h, e->spins, e->fields, e->energy, e->n, e->n_pad,
                                        e->R, e->R / e->n_models, st));
        e->launches += 1;
        return SG_OK;
    }
    if (e->dig && !getenv("SG_K2_SIMT")) {
        const size_t need = sg::fields_tc_spin_tiles_bytes(e->n, e->R);
        if (need > e->spin_tiles_cap) {
            SG_CUDA(cudaStreamSynchronize(st));
            cudaFree(e->spin_tiles);
            e->spin_tiles = nullptr;
            e->spin_tiles_cap = 0;
            cudaError_t ce = cudaMalloc(&e->spin_tiles, need);
            if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(spin tiles)", ce);
            e->spin_tiles_cap = need;
        }
        SG_CUDA(sg::launch_fields_tc(e->spins, e->n_pad, e->dig, e->scale, e->h, e->n, e->n_tc, e->R,
                                     e->spin_tiles, e->fields, e->n_pad, st));
        e->launches += 2;
    } else {
        SG_CUDA(sg::launch_fields(e->spins, e->n_pad, e->Jt, e->h, e->n, e->n_pad, e->R, e->fields,
                                  e->n_pad, st));
        e->launches += 1;
    }
    SG_CUDA(sg::launch_energies(e->spins, e->n_pad, e->fields, e->n_pad, e->h, e->n, e->R,
                                e->energy, st));
    e->launches += 1;
    return SG_OK;
}

int sg_init_fields(sg_engine* e, void* stream) {
    SG_REQUIRE(e && e->R > 0 && (e->Jt || e->csr || e->lat), "sg_init_fields: set model and replicas first");
    DeviceGuard g(e->device);
    int rc = compute_fields(e, static_cast<cudaStream_t>(stream));
    if (rc != SG_OK) return rc;
    e->fields_valid = true;
    return sg_reset_best(e, stream);
}

int sg_refresh_fields(sg_engine* e, void* stream) {
    SG_REQUIRE(e && e->R > 0 && e->fields_valid, "sg_refresh_fields: call sg_init_fields first");
    DeviceGuard g(e->device);
    return compute_fields(e, static_cast<cudaStream_t>(stream));
}

int sg_get_energies(sg_engine* e, float* energies, int on_device, void* stream) {
    SG_REQUIRE(e && energies && e->R > 0, "sg_get_energies: allocate replicas first");
    DeviceGuard g(e->device);
    return copy_out(energies, e->energy, (size_t)e->R * sizeof(float), on_device,
                    static_cast<cudaStream_t>(stream));
}

int sg_get_fields(sg_engine* e, float* fields, int on_device, void* stream) {
    SG_REQUIRE(e && fields && e->R > 0, "sg_get_fields: allocate replicas first");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t count = (size_t)e->R * e->n;
    if (e->lat) return fail(SG_ERR_UNSUPPORTED, "sg_get_fields: the lattice kernel keeps no local fields");
    if (e->csr) {
        if (on_device) {
            SG_CUDA(sg::launch_from_replica_minor_f32(e->c_fields, e->n, e->R, e->Rp, fields, st));
            e->launches++;
            return SG_OK;
        }
        float* tmpf = nullptr;
        int rcf;
        if ((rcf = dev_alloc(&tmpf, count)) != SG_OK) return rcf;
        cudaError_t cf = sg::launch_from_replica_minor_f32(e->c_fields, e->n, e->R, e->Rp, tmpf, st);
        e->launches++;
        if (cf == cudaSuccess)
            cf = cudaMemcpyAsync(fields, tmpf, count * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (cf == cudaSuccess) cf = cudaStreamSynchronize(st);
        cudaFree(tmpf);
        if (cf != cudaSuccess) return fail(SG_ERR_CUDA, "sg_get_fields (csr)", cf);
        return SG_OK;
    }
    if (on_device) {
        SG_CUDA(sg::launch_unpad_f32(e->fields, e->n_pad, fields, e->n, e->R, st));
        e->launches++;
        return SG_OK;
    }
    float* tmp = nullptr;
    int rc;
    if ((rc = dev_alloc(&tmp, count)) != SG_OK) return rc;
    cudaError_t ce = sg::launch_unpad_f32(e->fields, e->n_pad, tmp, e->n, e->R, st);
    e->launches++;
    if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(fields, tmp, count * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "sg_get_fields", ce);
    return SG_OK;
}

int sg_get_accepted(sg_engine* e, uint64_t* accepted, int on_device, void* stream) {
    SG_REQUIRE(e && accepted && e->R > 0, "sg_get_accepted: allocate replicas first");
    DeviceGuard g(e->device);
    return copy_out(accepted, e->accepted, (size_t)e->R * sizeof(uint64_t), on_device,
                    static_cast<cudaStream_t>(stream));
}

int sg_get_best(sg_engine* e, float* best_energy, int8_t* best_spins, int on_device,
                void* stream) {
    SG_REQUIRE(e && e->R > 0, "sg_get_best: allocate replicas first");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = SG_OK;
    if (best_energy)
        rc = copy_out(best_energy, e->best_energy, (size_t)e->R * sizeof(float), on_device, st);
    if (rc == SG_OK && best_spins)
        rc = e->csr   ? csr_get_i8(e, e->c_best, best_spins, on_device, st)
             : e->lat ? lat_get_spins(e, e->l_best, best_spins, on_device, st)
                      : get_unpadded_i8(e, e->best_spins, best_spins, on_device, st);
    return rc;
}

// ---- checkpoint restore: the state a resumed run needs besides the spins (sg_set_spins)
int sg_set_best(sg_engine* e, const float* best_energy, const int8_t* best_spins, int on_device, void* stream) {
    SG_REQUIRE(e && best_energy && best_spins && e->R > 0, "sg_set_best: allocate replicas first");
    SG_REQUIRE(!e->csr && !e->lat, "sg_set_best: dense models only");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)e->R * e->n;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    SG_CUDA(cudaMemcpyAsync(e->best_energy, best_energy, (size_t)e->R * sizeof(float), kind, st));
    const int8_t* src = best_spins;
    int8_t* tmp = nullptr;
    if (!on_device) {
        int rc = dev_alloc(&tmp, bytes);
        if (rc != SG_OK) return rc;
        SG_CUDA(cudaMemcpyAsync(tmp, best_spins, bytes, cudaMemcpyHostToDevice, st));
        src = tmp;
    }
    SG_CUDA(sg::launch_pad_spins(src, e->n, e->best_spins, e->n_pad, e->R, st));
    e->launches++;
    if (tmp) {
        SG_CUDA(cudaStreamSynchronize(st));
        cudaFree(tmp);
    } else if (!on_device) {
        SG_CUDA(cudaStreamSynchronize(st));
    }
    return SG_OK;
}

int sg_set_accepted(sg_engine* e, const uint64_t* accepted, int on_device, void* stream) {
    SG_REQUIRE(e && accepted && e->R > 0 && e->accepted, "sg_set_accepted: allocate replicas first");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SG_CUDA(cudaMemcpyAsync(e->accepted, accepted, (size_t)e->R * sizeof(uint64_t),
                            on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (!on_device) SG_CUDA(cudaStreamSynchronize(st));
    return SG_OK;
}

int sg_set_ladder_state(sg_engine* e, const int32_t* replica_at_rung, const uint32_t* attempts,
                        const uint32_t* accepts, int on_device, void* stream) {
    SG_REQUIRE(e && replica_at_rung && e->K > 0, "sg_set_ladder_state: call sg_set_ladder first");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t nstat = (size_t)e->L * (e->K > 1 ? e->K - 1 : 1);
    SG_CUDA(cudaMemcpyAsync(e->rep_at, replica_at_rung, (size_t)e->n_global * sizeof(int), kind, st));
    if (attempts) SG_CUDA(cudaMemcpyAsync(e->attempts, attempts, nstat * sizeof(unsigned int), kind, st));
    if (accepts) SG_CUDA(cudaMemcpyAsync(e->accepts, accepts, nstat * sizeof(unsigned int), kind, st));
    SG_CUDA(sg::launch_ladder_temps(e->rep_at, e->ladder, e->rep_temp, e->n_global, e->K, e->rep_lo, e->R, st));
    e->launches++;
    if (!on_device) SG_CUDA(cudaStreamSynchronize(st));
    return SG_OK;
}

int sg_upload_spins_async(sg_engine* e, const int8_t* host_spins, int slot, void* stream) {
    SG_REQUIRE(e && host_spins && e->R > 0, "sg_upload_spins_async: allocate replicas first");
    SG_REQUIRE(slot == 0 || slot == 1, "sg_upload_spins_async: slot must be 0 or 1");
    SG_REQUIRE(!e->csr && !e->lat, "sg_upload_spins_async: dense models only (use sg_set_spins)");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)e->R * e->n;
    if (e->up_cap[slot] < bytes) {
        SG_CUDA(cudaDeviceSynchronize());   // the old buffer may be in use on some stream
        cudaFree(e->up_stage[slot]);
        e->up_stage[slot] = nullptr;
        e->up_cap[slot] = 0;
        int rc = dev_alloc(&e->up_stage[slot], bytes);
        if (rc != SG_OK) return rc;
        e->up_cap[slot] = bytes;
    }
    SG_CUDA(cudaMemcpyAsync(e->up_stage[slot], host_spins, bytes, cudaMemcpyHostToDevice, st));
    return SG_OK;
}

int sg_set_spins_staged(sg_engine* e, int slot, void* stream) {
    SG_REQUIRE(e && e->R > 0 && (slot == 0 || slot == 1), "sg_set_spins_staged: bad argument");
    SG_REQUIRE(e->up_stage[slot] && e->up_cap[slot] >= (size_t)e->R * e->n,
               "sg_set_spins_staged: nothing was uploaded into this slot");
    DeviceGuard g(e->device);
    SG_CUDA(sg::launch_pad_spins(e->up_stage[slot], e->n, e->spins, e->n_pad, e->R,
                                 static_cast<cudaStream_t>(stream)));
    e->launches++;
    e->fields_valid = false;
    return SG_OK;
}

int sg_get_best_config(sg_engine* e, float* best_energy, int32_t* replica, int8_t* spins, int on_device,
                       void* stream) {
    SG_REQUIRE(e && e->R > 0 && e->fields_valid, "sg_get_best_config: call sg_init_fields first");
    SG_REQUIRE(!e->csr && !e->lat, "sg_get_best_config: dense models only (use sg_get_best)");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t need = 16 + (size_t)e->n;
    if (e->best_out_cap < need) {
        SG_CUDA(cudaStreamSynchronize(st));
        cudaFree(e->best_out);
        e->best_out = nullptr;
        e->best_out_cap = 0;
        int rc = dev_alloc(&e->best_out, need);
        if (rc != SG_OK) return rc;
        e->best_out_cap = need;
    }
    float* d_e = reinterpret_cast<float*>(e->best_out);
    int* d_i = reinterpret_cast<int*>(e->best_out + 4);
    int8_t* d_s = reinterpret_cast<int8_t*>(e->best_out + 16);
    SG_CUDA(sg::launch_best_config(e->best_energy, e->R, e->best_spins, e->n, e->n_pad, d_e, d_i, d_s, st));
    e->launches++;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (best_energy) SG_CUDA(cudaMemcpyAsync(best_energy, d_e, sizeof(float), kind, st));
    if (replica) SG_CUDA(cudaMemcpyAsync(replica, d_i, sizeof(int), kind, st));
    if (spins) SG_CUDA(cudaMemcpyAsync(spins, d_s, (size_t)e->n, kind, st));
    return SG_OK;
}

int sg_sweep(sg_engine* e, const sg_sweep_params* p, void* stream) {
    SG_REQUIRE(e && p, "sg_sweep: NULL argument");
    SG_REQUIRE(p->struct_size == sizeof(sg_sweep_params), "sg_sweep: struct_size mismatch");
    SG_REQUIRE(e->R > 0 && e->fields_valid, "sg_sweep: call sg_init_fields first");
    SG_REQUIRE(p->n_sweeps >= 0, "sg_sweep: n_sweeps < 0");
    SG_REQUIRE(p->rule != SG_RULE_WOLFF, "sg_sweep: the Wolff cluster move has its own entry point, sg_sweep_wolff");
    SG_REQUIRE(p->rule >= 0 && p->rule <= 2, "sg_sweep: unknown rule");
    SG_REQUIRE(p->rng_mode == SG_RNG_PHILOX || p->rng_mode == SG_RNG_INJECTED,
               "sg_sweep: unknown rng_mode");
    SG_REQUIRE(p->site_mode >= 0 && p->site_mode <= 4, "sg_sweep: unknown site_mode");
    SG_REQUIRE(p->site_mode != SG_SITES_CHECKERBOARD || (e->lat && !e->grp),
               "sg_sweep: SG_SITES_CHECKERBOARD is the order of lattice models only");
    SG_REQUIRE(p->site_mode != SG_SITES_EXPLICIT || p->sites, "sg_sweep: explicit sites missing");
    SG_REQUIRE(p->rng_mode != SG_RNG_INJECTED || p->uniforms, "sg_sweep: injected uniforms missing");
    SG_REQUIRE(p->temps || e->rep_temp, "sg_sweep: no temperatures (pass temps or set a ladder)");
    if (p->n_sweeps == 0) return SG_OK;
    SG_REQUIRE((long long)p->n_sweeps * e->n < (1LL << 31) - 64,
               "sg_sweep: n_sweeps * n must stay below 2^31 per launch (cut the run into launches)");
    DeviceGuard g(e->device);
    const int gmax = sg::sweep_max_replicas_per_block(e->n_pad);
    int G = p->replicas_per_block;
    SG_REQUIRE(G >= 0 && G <= gmax, "sg_sweep: replicas_per_block out of range");
    if (G == 0) {
        // fewest waves first, then the fewest replicas per block that achieves it
        const long long sms = e->sm_count;
        long long best_cost = -1;
        for (int cand = 1; cand <= gmax; ++cand) {
            const long long blocks = (e->R + cand - 1) / cand;
            const long long waves = (blocks + sms - 1) / sms;
            const long long cost = waves * (gmax + 2LL * cand);  // per-attempt cost ~ a + b*G
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                G = cand;
            }
        }
    }
    sg::SweepDev a{};
    a.Jt = e->Jt;
    a.h = e->h;
    a.spins = e->spins;
    a.fields = e->fields;
    a.energy = e->energy;
    a.best_energy = e->best_energy;
    a.best_spins = e->best_spins;
    a.accepted = e->accepted;
    a.energy_trace = p->energy_trace;
    if (p->temps) {
        a.temps = p->temps;
        a.t_ss = p->temps_sweep_stride;
        a.t_rs = p->temps_replica_stride;
    } else {
        a.temps = e->rep_temp;
        a.t_ss = 0;
        a.t_rs = 1;
    }
    a.sites = p->sites;
    a.s_bs = p->sites_block_stride;
    a.s_ss = p->sites_sweep_stride;
    a.uniforms = p->uniforms;
    a.seed = p->seed;
    a.sweep_base = p->sweep_base;
    a.n = e->n;
    a.n_pad = e->n_pad;
    a.R = e->R;
    a.G = G;
    a.n_sweeps = p->n_sweeps;
    a.rule = p->rule;
    a.site_mode = p->site_mode;
    a.track_best = p->track_best ? 1 : 0;
    a.dbg = e->dbg;
    a.site_de = p->site_energy_changes;
    a.rpm = (e->stacked && !e->csr && !e->lat) ? e->R / e->n_models : 0;
    SG_REQUIRE(p->replica_base >= 0 && (!e->lat || p->replica_base % 32 == 0),
               "sg_sweep: replica_base must be >= 0 (a multiple of 32 for lattice models)");
    a.rep_base = p->replica_base;
    SG_REQUIRE(!a.site_de || (!e->csr && !e->lat && p->kernel != SG_KERNEL_TC &&
                              (p->kernel == SG_KERNEL_SIMT || p->kernel == SG_KERNEL_SMALL ||
                               p->rng_mode == SG_RNG_INJECTED || sg::sweep_small_supported(e->n))),
               "sg_sweep: site_energy_changes needs a dense model on a sequential-FMA kernel");
    if (e->csr) return csr_sweep(e, p, a, static_cast<cudaStream_t>(stream));
    if (e->lat) return lat_sweep(e, p, a, static_cast<cudaStream_t>(stream));
    const bool inject = (p->rng_mode == SG_RNG_INJECTED);
    SG_REQUIRE(p->kernel >= SG_KERNEL_AUTO && p->kernel <= SG_KERNEL_SMALL, "sg_sweep: unknown kernel");
    const bool shared_order = p->site_mode != SG_SITES_RANDOM_PER_BLOCK &&
                              !(p->site_mode == SG_SITES_EXPLICIT && p->sites_block_stride != 0) &&
                              p->replicas_per_block == 0;
    const bool small_ok = e->Jt && sg::sweep_small_supported(e->n) && shared_order;
    if (e->stacked)
        SG_REQUIRE(small_ok && (p->kernel == SG_KERNEL_AUTO || p->kernel == SG_KERNEL_SMALL),
                   "sg_sweep: stacked models run on the small-model kernel (one site order for the "
                   "grid, replicas_per_block = 0)");
    if (p->kernel == SG_KERNEL_SMALL)
        SG_REQUIRE(small_ok, "sg_sweep: the small-model kernel needs a dense model with n <= 224, one "
                             "site order for the grid and replicas_per_block = 0");
    if (p->kernel == SG_KERNEL_SMALL || (p->kernel == SG_KERNEL_AUTO && small_ok)) {
        const size_t need = sg::csr_sites_bytes(e->n, p->n_sweeps);
        if (need > e->c_sites_cap) {
            SG_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
            cudaFree(e->c_sites);
            e->c_sites = nullptr;
            e->c_sites_cap = 0;
            cudaError_t ce = cudaMalloc(&e->c_sites, need);
            if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(site tables)", ce);
            e->c_sites_cap = need;
        }
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        SG_CUDA(sg::launch_sites_table(a, static_cast<int*>(e->c_sites), st));
        if (e->profiling) e->timer.begin(0, st);
        SG_CUDA(sg::launch_sweep_small(a, p->rng_mode == SG_RNG_INJECTED,
                                       static_cast<const int*>(e->c_sites), st));
        if (e->profiling) e->timer.end(st);
        e->launches += 2;
        return SG_OK;
    }
    SG_REQUIRE(p->coupling_planes >= 0 && p->coupling_planes <= 3,
               "sg_sweep: coupling_planes must be 0..3");
    const bool tc_ok = e->Jp && sg::sweep_tc_supported(e->n, e->n_tc) &&
                       p->site_mode != SG_SITES_RANDOM_PER_BLOCK &&
                       !(p->site_mode == SG_SITES_EXPLICIT && p->sites_block_stride != 0) &&
                       p->replicas_per_block == 0;
    bool use_tc = false;
    if (p->kernel == SG_KERNEL_TC) {
        SG_REQUIRE(tc_ok, "sg_sweep: the tensor-core kernel needs a dense model with 16 <= n <= 4096, "
                          "one site order for the grid and replicas_per_block = 0");
        use_tc = true;
    } else if (p->kernel == SG_KERNEL_AUTO) {
        // replay (injected uniforms) stays on the sequential-FMA kernel unless asked otherwise
        use_tc = tc_ok && !inject;
    }
    if (use_tc) {
        const size_t need = sg::sweep_tc_sites_bytes(e->n, p->n_sweeps, e->R);
        if (need > e->tc_sites_cap) {
            // the previous table may still be in use by a launch in flight on this stream
            SG_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
            cudaFree(e->tc_sites);
            e->tc_sites = nullptr;
            e->tc_sites_cap = 0;
            cudaError_t ce = cudaMalloc(&e->tc_sites, need);
            if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(site tables)", ce);
            e->tc_sites_cap = need;
        }
        // coupling_planes = 0: as many planes as the couplings need to be exact -- three for
        // arbitrary fp32 values, one when every coupling is a bf16 value (integer couplings up to
        // 256): a third of the tensor-core instructions and of the operand stream
        if (p->coupling_planes == 0 && e->planes_needed == 0) {
            int used[2] = {1, 1};
            if (e->plane_flags) {
                SG_CUDA(cudaMemcpyAsync(used, e->plane_flags, sizeof(used), cudaMemcpyDeviceToHost,
                                        static_cast<cudaStream_t>(stream)));
                SG_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
            }
            e->planes_needed = used[1] ? 3 : used[0] ? 2 : 1;
        }
        const int planes = p->coupling_planes ? p->coupling_planes : e->planes_needed;
        const size_t per_sweep = sg::sweep_tc_stream_bytes_per_sweep(e->n, e->n_tc, planes);
        size_t want = per_sweep * (size_t)p->n_sweeps;
        const size_t cap_max = (size_t)1 << 30;
        if (want > cap_max) want = (cap_max / per_sweep ? cap_max / per_sweep : 1) * per_sweep;
        if (want > e->tc_stream_cap) {
            SG_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
            cudaFree(e->tc_stream);
            e->tc_stream = nullptr;
            e->tc_stream_cap = 0;
            cudaError_t ce = cudaMalloc(&e->tc_stream, want);
            if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(operand stream)", ce);
            e->tc_stream_cap = want;
        }
        a.G = 16;
        SG_CUDA(sg::launch_sweep_tc(a, e->Jp, e->n_tc, planes, inject, e->tc_sites, e->tc_stream,
                                    e->tc_stream_cap, &e->launches,
                                    e->profiling ? &e->timer : nullptr,
                                    static_cast<cudaStream_t>(stream)));
        return SG_OK;
    }
    const int grid = (e->R + G - 1) / G;
    if (e->profiling) e->timer.begin(0, static_cast<cudaStream_t>(stream));
    SG_CUDA(sg::launch_sweep(a, inject, grid, static_cast<cudaStream_t>(stream)));
    if (e->profiling) e->timer.end(static_cast<cudaStream_t>(stream));
    e->launches++;
    return SG_OK;
}

int sg_sweep_wolff(sg_engine* e, const sg_wolff_params* p, void* stream) {
    SG_REQUIRE(e && p, "sg_sweep_wolff: NULL argument");
    SG_REQUIRE(p->struct_size == sizeof(sg_wolff_params), "sg_sweep_wolff: struct_size mismatch");
    if (e->csr || e->lat || e->stacked || !e->Jt)
        return fail(SG_ERR_UNSUPPORTED, "sg_sweep_wolff: the cluster move takes one dense model (sg_set_model_dense)");
    SG_REQUIRE(e->R > 0 && e->fields_valid, "sg_sweep_wolff: call sg_init_fields first");
    SG_REQUIRE(p->n_sweeps >= 0, "sg_sweep_wolff: n_sweeps < 0");
    SG_REQUIRE(p->rng_mode == SG_RNG_PHILOX || p->rng_mode == SG_RNG_INJECTED, "sg_sweep_wolff: unknown rng_mode");
    SG_REQUIRE(p->site_mode == SG_SITES_SEQUENTIAL || p->site_mode == SG_SITES_RANDOM ||
                   p->site_mode == SG_SITES_EXPLICIT,
               "sg_sweep_wolff: site_mode must be SEQUENTIAL, RANDOM or EXPLICIT");
    SG_REQUIRE(p->site_mode != SG_SITES_EXPLICIT || p->sites, "sg_sweep_wolff: explicit sites missing");
    const bool inject = p->rng_mode == SG_RNG_INJECTED;
    SG_REQUIRE(!inject || (p->uniforms && p->cursor && p->uniforms_per_replica >= 0),
               "sg_sweep_wolff: injected mode needs uniforms, cursor and uniforms_per_replica");
    SG_REQUIRE(p->temps || e->rep_temp, "sg_sweep_wolff: no temperatures (pass temps or set a ladder)");
    SG_REQUIRE(p->replica_base >= 0, "sg_sweep_wolff: replica_base must be >= 0");
    if (p->n_sweeps == 0) return SG_OK;
    SG_REQUIRE((long long)p->n_sweeps * e->n < (1LL << 31) - 64,
               "sg_sweep_wolff: n_sweeps * n must stay below 2^31 per call");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (!e->Jrow) {   // Jt[i][j] = J[j][i]; the growth reads rows of J
        if ((rc = dev_alloc(&e->Jrow, (size_t)e->n * e->n_pad)) != SG_OK) return rc;
        SG_CUDA(sg::launch_pad_transpose(e->Jt, e->n_pad, e->n, e->Jrow, e->n_pad, st));
        e->launches++;
    }
    if (!e->wolff_status) {
        if ((rc = dev_alloc(&e->wolff_status, (size_t)1)) != SG_OK) return rc;
    }
    if (e->wolff_max_deg < 0) {   // once per model: do the rows' negative couplings fit neighbour lists?
        SG_CUDA(cudaMemsetAsync(e->wolff_status, 0, sizeof(int), st));
        SG_CUDA(sg::launch_wolff_neg_count(e->Jrow, e->n, e->n_pad, e->wolff_status, st));
        int max_deg = 0;
        SG_CUDA(cudaMemcpyAsync(&max_deg, e->wolff_status, sizeof(int), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaStreamSynchronize(st));
        e->wolff_max_deg = max_deg;
        e->launches++;
        if (max_deg <= 32 && e->n < 65535) {
            if ((rc = dev_alloc(&e->wolff_nb_col, (size_t)e->n * 32)) != SG_OK) return rc;
            if ((rc = dev_alloc(&e->wolff_nb_val, (size_t)e->n * 32)) != SG_OK) return rc;
            SG_CUDA(sg::launch_wolff_neg_fill(e->Jrow, e->n, e->n_pad, e->wolff_nb_col, e->wolff_nb_val, st));
            e->launches++;
        }
    }
    if (inject) SG_CUDA(cudaMemsetAsync(e->wolff_status, 0, sizeof(int), st));
    sg::WolffDev a{};
    a.Jrow = e->Jrow;
    a.nb_col = e->wolff_nb_col;
    a.nb_val = e->wolff_nb_val;
    a.spins = e->spins;
    a.accepted = e->accepted;
    if (p->temps) {
        a.temps = p->temps;
        a.t_ss = p->temps_sweep_stride;
        a.t_rs = p->temps_replica_stride;
    } else {
        a.temps = e->rep_temp;
        a.t_ss = 0;
        a.t_rs = 1;
    }
    if (p->site_mode == SG_SITES_EXPLICIT) {
        a.sites = p->sites;
        a.s_rs = p->sites_replica_stride;
        a.s_ss = p->sites_sweep_stride;
    } else {   // one list per sweep for all replicas, from the same Philox stream as the other kernels
        const size_t need = sg::csr_sites_bytes(e->n, p->n_sweeps);
        if (need > e->c_sites_cap) {
            SG_CUDA(cudaStreamSynchronize(st));
            cudaFree(e->c_sites);
            e->c_sites = nullptr;
            e->c_sites_cap = 0;
            cudaError_t ce = cudaMalloc(&e->c_sites, need);
            if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(site tables)", ce);
            e->c_sites_cap = need;
        }
        sg::SweepDev t{};
        t.site_mode = p->site_mode;
        t.seed = p->seed;
        t.sweep_base = p->sweep_base;
        t.n = e->n;
        t.n_sweeps = p->n_sweeps;
        SG_CUDA(sg::launch_sites_table(t, static_cast<int*>(e->c_sites), st));
        e->launches++;
        a.sites = static_cast<const int*>(e->c_sites);
        a.s_rs = 0;
        a.s_ss = e->n;
    }
    a.uniforms = p->uniforms;
    a.u_rs = p->uniforms_replica_stride;
    a.u_len = p->uniforms_per_replica;
    a.cursor = reinterpret_cast<long long*>(p->cursor);
    a.status = e->wolff_status;
    a.seed = p->seed;
    a.n = e->n;
    a.n_pad = e->n_pad;
    a.R = e->R;
    a.rep_base = p->replica_base;
    for (int s = 0; s < p->n_sweeps; ++s) {
        a.sweep = s;
        a.sweep_abs = p->sweep_base + (unsigned long long)s;
        if (e->profiling) e->timer.begin(0, st);
        SG_CUDA(sg::launch_wolff(a, inject, st));
        if (e->profiling) e->timer.end(st);
        e->launches++;
        if ((rc = compute_fields(e, st)) != SG_OK) return rc;   // exact fields / energies
        if (p->energy_trace || p->track_best) {
            SG_CUDA(sg::launch_wolff_record(e->energy, e->best_energy, e->spins, e->best_spins,
                                            p->energy_trace ? p->energy_trace + (size_t)s * e->R : nullptr,
                                            e->n_pad, e->R, p->track_best ? 1 : 0, st));
            e->launches++;
        }
    }
    if (inject) {
        int dry = 0;
        SG_CUDA(cudaMemcpyAsync(&dry, e->wolff_status, sizeof(int), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaStreamSynchronize(st));
        if (dry) return fail(SG_ERR_INVALID, "sg_sweep_wolff: a replica needed more uniforms than uniforms_per_replica");
    }
    return SG_OK;
}

int sg_set_ladder(sg_engine* e, int n_rungs, const double* ladder_temps, void* stream) {
    SG_REQUIRE(e && e->R > 0, "sg_set_ladder: allocate replicas first");
    return sg_set_ladder_sharded(e, n_rungs, ladder_temps, e->R, 0, stream);
}

int sg_set_ladder_sharded(sg_engine* e, int n_rungs, const double* ladder_temps, int n_global_replicas,
                          int replica_offset, void* stream) {
    SG_REQUIRE(e && ladder_temps && e->R > 0, "sg_set_ladder: allocate replicas first");
    SG_REQUIRE(n_rungs >= 1 && n_global_replicas >= e->R && n_global_replicas % n_rungs == 0,
               "sg_set_ladder: the (global) replica count must be a multiple of n_rungs");
    SG_REQUIRE(replica_offset >= 0 && replica_offset + e->R <= n_global_replicas,
               "sg_set_ladder_sharded: local replicas must lie inside the global range");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    free_ladder(e);
    const int K = n_rungs, L = n_global_replicas / n_rungs;
    int rc;
    if ((rc = dev_alloc(&e->rep_at, (size_t)n_global_replicas)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->rep_temp, (size_t)e->R)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->ladder, (size_t)K)) != SG_OK) return rc;
    const size_t nstat = (size_t)L * (K > 1 ? K - 1 : 1);
    if ((rc = dev_alloc(&e->attempts, nstat)) != SG_OK) return rc;
    if ((rc = dev_alloc(&e->accepts, nstat)) != SG_OK) return rc;
    SG_CUDA(cudaMemcpyAsync(e->ladder, ladder_temps, (size_t)K * sizeof(double),
                            cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemsetAsync(e->attempts, 0, nstat * sizeof(unsigned int), st));
    SG_CUDA(cudaMemsetAsync(e->accepts, 0, nstat * sizeof(unsigned int), st));
    SG_CUDA(sg::launch_ladder_init(e->rep_at, e->rep_temp, e->ladder, n_global_replicas, K,
                                   replica_offset, e->R, st));
    e->launches++;
    SG_CUDA(cudaStreamSynchronize(st));  // ladder_temps is a host buffer
    e->K = K;
    e->L = L;
    e->n_global = n_global_replicas;
    e->rep_lo = replica_offset;
    return SG_OK;
}

int sg_exchange(sg_engine* e, const sg_exchange_params* p, void* stream) {
    SG_REQUIRE(e && p, "sg_exchange: NULL argument");
    SG_REQUIRE(p->struct_size == sizeof(sg_exchange_params), "sg_exchange: struct_size mismatch");
    SG_REQUIRE(e->K > 0, "sg_exchange: call sg_set_ladder first");
    SG_REQUIRE(p->parity == 0 || p->parity == 1, "sg_exchange: parity must be 0 or 1");
    SG_REQUIRE(p->method == SG_EXCHANGE_NEAREST || p->method == SG_EXCHANGE_ALL_PAIRS,
               "sg_exchange: unknown method");
    SG_REQUIRE(p->rng_mode != SG_RNG_INJECTED || p->uniforms, "sg_exchange: uniforms missing");
    SG_REQUIRE(p->energies_all || e->n_global == e->R,
               "sg_exchange: a sharded ladder needs the all-gathered energy table (energies_all)");
    DeviceGuard g(e->device);
    sg::ExchangeDev a{};
    a.rep_at = e->rep_at;
    a.rep_temp = e->rep_temp;
    a.ladder = e->ladder;
    a.energy = p->energies_all ? p->energies_all : e->energy;
    a.attempts = e->attempts;
    a.accepts = e->accepts;
    a.uniforms = p->uniforms;
    a.seed = p->seed;
    a.round = p->round;
    a.L = e->L;
    a.K = e->K;
    a.parity = p->parity;
    a.inject = (p->rng_mode == SG_RNG_INJECTED);
    a.method = p->method;
    a.rep_lo = e->rep_lo;
    a.rep_n = e->R;
    SG_CUDA(sg::launch_exchange(a, static_cast<cudaStream_t>(stream)));
    e->launches++;
    return SG_OK;
}

int sg_check_target(sg_engine* e, int which, float target, int32_t round, int32_t* hit, void* stream) {
    SG_REQUIRE(e && hit && e->R > 0 && e->fields_valid, "sg_check_target: call sg_init_fields first");
    SG_REQUIRE(which == 0 || which == 1, "sg_check_target: which must be 0 (current) or 1 (best so far)");
    DeviceGuard g(e->device);
    SG_CUDA(sg::launch_check_target(which ? e->best_energy : e->energy, e->R, e->rep_lo, target, round,
                                    hit, static_cast<cudaStream_t>(stream)));
    e->launches++;
    return SG_OK;
}

int sg_adaptive_temperature(sg_engine* e, int replica, uint64_t accepted_base, int sweep, int window,
                            double target_acceptance, double adaptation_rate, double final_temp,
                            const double* base_temps, double* state, double* temps_out, void* stream) {
    SG_REQUIRE(e && base_temps && state && temps_out, "sg_adaptive_temperature: NULL argument");
    SG_REQUIRE(e->R > 0 && replica >= 0 && replica < e->R && e->accepted,
               "sg_adaptive_temperature: allocate replicas first");
    SG_REQUIRE(sweep >= 0 && window >= 1, "sg_adaptive_temperature: bad sweep / window");
    DeviceGuard g(e->device);
    SG_CUDA(sg::launch_adaptive_temperature(e->accepted + replica, accepted_base, e->n, sweep, window,
                                            target_acceptance, adaptation_rate, final_temp, base_temps,
                                            state, temps_out, static_cast<cudaStream_t>(stream)));
    e->launches++;
    return SG_OK;
}

int sg_exchange_chain(int device, void* rows, int64_t row_stride_bytes, int64_t row_bytes,
                      int n_replicas, float* energies, const float* temperatures,
                      const float* uniforms, uint64_t seed, uint64_t round, int32_t* n_accepted,
                      void* stream) {
    SG_REQUIRE(rows && energies && temperatures && n_accepted, "sg_exchange_chain: NULL argument");
    SG_REQUIRE(n_replicas >= 1 && row_bytes >= 1 && row_stride_bytes >= row_bytes,
               "sg_exchange_chain: bad shape");
    *n_accepted = 0;
    if (n_replicas == 1) return SG_OK;
    DeviceGuard g(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t head = sg::exchange_chain_header_bytes(n_replicas);
    void* scratch = nullptr;
    cudaError_t ce = cudaMallocAsync(&scratch, head + (size_t)n_replicas * (size_t)row_bytes, st);
    if (ce != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMallocAsync(exchange scratch)", ce);
    ce = sg::launch_exchange_chain(rows, row_stride_bytes, row_bytes, n_replicas, energies,
                                   temperatures, uniforms, seed, round, scratch, st);
    if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(n_accepted, static_cast<int*>(scratch) + n_replicas, sizeof(int),
                             cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(scratch, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "sg_exchange_chain", ce);
    return SG_OK;
}

int sg_get_ladder_state(sg_engine* e, int32_t* replica_at_rung, double* replica_temps,
                        uint32_t* attempts, uint32_t* accepts, int on_device, void* stream) {
    SG_REQUIRE(e && e->K > 0, "sg_get_ladder_state: call sg_set_ladder first");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t nstat = (size_t)e->L * (e->K > 1 ? e->K - 1 : 1);
    int rc = SG_OK;
    if (replica_at_rung)
        rc = copy_out(replica_at_rung, e->rep_at, (size_t)e->n_global * sizeof(int), on_device, st);
    if (rc == SG_OK && replica_temps)
        rc = copy_out(replica_temps, e->rep_temp, (size_t)e->R * sizeof(double), on_device, st);
    if (rc == SG_OK && attempts)
        rc = copy_out(attempts, e->attempts, nstat * sizeof(unsigned int), on_device, st);
    if (rc == SG_OK && accepts)
        rc = copy_out(accepts, e->accepts, nstat * sizeof(unsigned int), on_device, st);
    return rc;
}

int sg_batch_energies(sg_engine* e, int batch, const int8_t* spins, float* energies, float* fields,
                      int on_device, void* stream) {
    SG_REQUIRE(e && spins && (e->Jt || e->csr || e->lat), "sg_batch_energies: set the model first");
    SG_REQUIRE(batch >= 1, "sg_batch_energies: batch < 1");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (e->csr) return csr_batch_energies(e, batch, spins, energies, fields, on_device, st);
    if (e->lat) return lat_batch_energies(e, batch, spins, energies, fields, on_device, st);
    const size_t B = (size_t)batch, n = (size_t)e->n, np = (size_t)e->n_pad;
    if (e->stacked) {
        // stacked models: configuration b belongs to model b / (batch / n_models)
        SG_REQUIRE(batch % e->n_models == 0, "sg_batch_energies: stacked models need a multiple of "
                                             "n_models configurations (model-major)");
        auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
        const size_t o_sp = 0, o_fp = o_sp + up16(B * np), o_ed = o_fp + up16(B * np * 4),
                     o_in = o_ed + up16(B * 4), o_fo = o_in + up16(B * n), need = o_fo + up16(B * n * 4);
        if (need > e->stage_cap) {
            SG_CUDA(cudaStreamSynchronize(st));
            cudaFree(e->stage); e->stage = nullptr; e->stage_cap = 0;
            cudaError_t ce0 = cudaMalloc(&e->stage, need);
            if (ce0 != cudaSuccess) return fail(SG_ERR_NOMEM, "cudaMalloc(batch staging)", ce0);
            e->stage_cap = need;
        }
        unsigned char* base = static_cast<unsigned char*>(e->stage);
        int8_t* sp = reinterpret_cast<int8_t*>(base + o_sp);
        float* fp = reinterpret_cast<float*>(base + o_fp);
        float* ed = reinterpret_cast<float*>(base + o_ed);
        int8_t* sin = reinterpret_cast<int8_t*>(base + o_in);
        float* fo = reinterpret_cast<float*>(base + o_fo);
        const int8_t* src = spins;
        if (!on_device) {
            SG_CUDA(cudaMemcpyAsync(sin, spins, B * n, cudaMemcpyHostToDevice, st));
            src = sin;
        }
        SG_CUDA(sg::launch_pad_spins(src, e->n, sp, e->n_pad, batch, st));
        SG_CUDA(sg::launch_fields_small(e->Jt, e->h, sp, fp, ed, e->n, e->n_pad, batch,
                                        batch / e->n_models, st));
        e->launches += 2;
        if (energies)
            SG_CUDA(cudaMemcpyAsync(energies, ed, B * sizeof(float),
                                    on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        if (fields) {
            float* dstf = on_device ? fields : fo;
            SG_CUDA(sg::launch_unpad_f32(fp, e->n_pad, dstf, e->n, batch, st));
            if (!on_device)
                SG_CUDA(cudaMemcpyAsync(fields, fo, B * n * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        if (!on_device) SG_CUDA(cudaStreamSynchronize(st));
        return SG_OK;
    }
    int8_t *s_in = nullptr, *s_pad = nullptr;
    unsigned char* tiles = nullptr;
    float *f_pad = nullptr, *e_dev = nullptr, *f_out = nullptr;
    int rc = SG_OK;
    cudaError_t ce = cudaSuccess;
    do {
        if ((rc = dev_alloc(&s_pad, B * np)) != SG_OK) break;
        if ((rc = dev_alloc(&f_pad, B * np)) != SG_OK) break;
        if ((rc = dev_alloc(&e_dev, B)) != SG_OK) break;
        const int8_t* src = spins;
        if (!on_device) {
            if ((rc = dev_alloc(&s_in, B * n)) != SG_OK) break;
            if ((ce = cudaMemcpyAsync(s_in, spins, B * n, cudaMemcpyHostToDevice, st))) break;
            src = s_in;
        }
        if ((ce = sg::launch_pad_spins(src, e->n, s_pad, e->n_pad, batch, st))) break;
        if (e->dig && !getenv("SG_K2_SIMT")) {
            if ((rc = dev_alloc(&tiles, sg::fields_tc_spin_tiles_bytes(e->n, batch))) != SG_OK) break;
            if ((ce = cudaMemsetAsync(f_pad, 0, B * np * sizeof(float), st))) break;
            if ((ce = sg::launch_fields_tc(s_pad, np, e->dig, e->scale, e->h, e->n, e->n_tc, batch,
                                           tiles, f_pad, np, st)))
                break;
        } else if ((ce = sg::launch_fields(s_pad, np, e->Jt, e->h, e->n, e->n_pad, batch, f_pad, np,
                                           st))) {
            break;
        }
        if ((ce = sg::launch_energies(s_pad, np, f_pad, np, e->h, e->n, batch, e_dev, st))) break;
        e->launches += 3;
        if (energies) {
            if ((ce = cudaMemcpyAsync(energies, e_dev, B * sizeof(float),
                                      on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                      st)))
                break;
        }
        if (fields) {
            float* dstf = fields;
            if (!on_device) {
                if ((rc = dev_alloc(&f_out, B * n)) != SG_OK) break;
                dstf = f_out;
            }
            if ((ce = sg::launch_unpad_f32(f_pad, e->n_pad, dstf, e->n, batch, st))) break;
            e->launches++;
            if (!on_device &&
                (ce = cudaMemcpyAsync(fields, f_out, B * n * sizeof(float), cudaMemcpyDeviceToHost,
                                      st)))
                break;
        }
        ce = cudaStreamSynchronize(st);  // scratch is freed below
    } while (0);
    cudaFree(s_in);
    cudaFree(s_pad);
    cudaFree(tiles);
    cudaFree(f_pad);
    cudaFree(e_dev);
    cudaFree(f_out);
    if (rc != SG_OK) return rc;
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "sg_batch_energies", ce);
    return SG_OK;
}

int sg_measure_stream_bandwidth(sg_engine* e, int64_t bytes, int iters, int stagger,
                                double* gbps_out) {
    SG_REQUIRE(e && gbps_out, "sg_measure_stream_bandwidth: NULL argument");
    SG_REQUIRE(bytes >= (1 << 20) && iters >= 1, "sg_measure_stream_bandwidth: bytes >= 1 MiB");
    DeviceGuard g(e->device);
    const int64_t n_vec = (bytes / 16) / 1024 * 1024;
    float4* buf = nullptr;
    float* sink = nullptr;
    int rc;
    if ((rc = dev_alloc(&buf, (size_t)n_vec)) != SG_OK) return rc;
    if ((rc = dev_alloc(&sink, 1)) != SG_OK) {
        cudaFree(buf);
        return rc;
    }
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    cudaError_t ce = cudaMemset(buf, 0, (size_t)n_vec * 16);
    const int grid = e->sm_count;  // one block per SM, like the sweep kernel
    const int sg_ = stagger ? 1 : 0;
    if (ce == cudaSuccess) ce = sg::launch_stream_probe(buf, n_vec, 1, sg_, sink, grid, 0);  // warm
    if (ce == cudaSuccess) ce = cudaEventRecord(t0, 0);
    if (ce == cudaSuccess) ce = sg::launch_stream_probe(buf, n_vec, iters, sg_, sink, grid, 0);
    if (ce == cudaSuccess) ce = cudaEventRecord(t1, 0);
    if (ce == cudaSuccess) ce = cudaEventSynchronize(t1);
    float ms = 0.0f;
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, t0, t1);
    e->launches += 2;
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(buf);
    cudaFree(sink);
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "sg_measure_stream_bandwidth", ce);
    *gbps_out = (double)n_vec * 16.0 * (double)iters * (double)grid / ((double)ms * 1.0e6);
    return SG_OK;
}

int sg_measure_tma_stream(sg_engine* e, int64_t bytes, int row_bytes, int depth, int n_rows,
                          int stagger, double* gbps_out) {
    SG_REQUIRE(e && gbps_out, "sg_measure_tma_stream: NULL argument");
    SG_REQUIRE(row_bytes >= 512 && row_bytes % 16 == 0 && depth >= 1 && n_rows >= 1 &&
                   (size_t)depth * row_bytes + 128 <= 227 * 1024 && bytes >= row_bytes,
               "sg_measure_tma_stream: bad shape");
    DeviceGuard g(e->device);
    const int64_t buf_rows = bytes / row_bytes;
    float* buf = nullptr;
    float* sink = nullptr;
    int rc;
    if ((rc = dev_alloc(&buf, (size_t)(buf_rows * row_bytes / 4))) != SG_OK) return rc;
    if ((rc = dev_alloc(&sink, 1)) != SG_OK) {
        cudaFree(buf);
        return rc;
    }
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    cudaError_t ce = cudaMemset(buf, 0, (size_t)buf_rows * row_bytes);
    const int grid = e->sm_count;
    if (ce == cudaSuccess)
        ce = sg::launch_tma_probe(buf, buf_rows, row_bytes, n_rows, depth, stagger, sink, grid, 0);
    if (ce == cudaSuccess) ce = cudaEventRecord(t0, 0);
    if (ce == cudaSuccess)
        ce = sg::launch_tma_probe(buf, buf_rows, row_bytes, n_rows, depth, stagger, sink, grid, 0);
    if (ce == cudaSuccess) ce = cudaEventRecord(t1, 0);
    if (ce == cudaSuccess) ce = cudaEventSynchronize(t1);
    float ms = 0.0f;
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, t0, t1);
    e->launches += 2;
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(buf);
    cudaFree(sink);
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "sg_measure_tma_stream", ce);
    *gbps_out = (double)row_bytes * n_rows * grid / ((double)ms * 1.0e6);
    return SG_OK;
}

int sg_tc_selftest(sg_engine* e, int planes, const int32_t* sites16, const float* deltas,
                   const float* fields_in, float* fields_out) {
    SG_REQUIRE(e && sites16 && deltas && fields_in && fields_out, "sg_tc_selftest: NULL argument");
    SG_REQUIRE(e->Jp, "sg_tc_selftest: no bf16 planes (set a dense model with n <= 4096)");
    SG_REQUIRE(planes >= 1 && planes <= 3, "sg_tc_selftest: planes must be 1..3");
    for (int k = 0; k < 16; ++k)
        SG_REQUIRE(sites16[k] >= 0 && sites16[k] < e->n, "sg_tc_selftest: site out of range");
    DeviceGuard g(e->device);
    const size_t nf = (size_t)16 * e->n_tc;
    int* d_sites = nullptr;
    float *d_delta = nullptr, *d_in = nullptr, *d_out = nullptr;
    int rc = SG_OK;
    cudaError_t ce = cudaSuccess;
    do {
        if ((rc = dev_alloc(&d_sites, 16)) != SG_OK) break;
        if ((rc = dev_alloc(&d_delta, 256)) != SG_OK) break;
        if ((rc = dev_alloc(&d_in, nf)) != SG_OK) break;
        if ((rc = dev_alloc(&d_out, nf)) != SG_OK) break;
        if ((ce = cudaMemcpy(d_sites, sites16, 16 * sizeof(int), cudaMemcpyHostToDevice))) break;
        if ((ce = cudaMemcpy(d_delta, deltas, 256 * sizeof(float), cudaMemcpyHostToDevice))) break;
        if ((ce = cudaMemset(d_in, 0, nf * sizeof(float)))) break;
        if ((ce = cudaMemcpy2D(d_in, (size_t)e->n_tc * 4, fields_in, (size_t)e->n * 4,
                               (size_t)e->n * 4, 16, cudaMemcpyHostToDevice)))
            break;
        if ((ce = sg::launch_tc_selftest(e->Jp, e->n, e->n_tc, planes, d_sites, d_delta, d_in, d_out,
                                         0)))
            break;
        e->launches++;
        if ((ce = cudaMemcpy2D(fields_out, (size_t)e->n * 4, d_out, (size_t)e->n_tc * 4,
                               (size_t)e->n * 4, 16, cudaMemcpyDeviceToHost)))
            break;
    } while (0);
    cudaFree(d_sites);
    cudaFree(d_delta);
    cudaFree(d_in);
    cudaFree(d_out);
    if (rc != SG_OK) return rc;
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "sg_tc_selftest", ce);
    return SG_OK;
}

int sg_set_profiling(sg_engine* e, int enable) {
    SG_REQUIRE(e, "sg_set_profiling: NULL engine");
    DeviceGuard g(e->device);
    double ms[2];
    unsigned long long cnt[2];
    e->timer.collect(ms, cnt);
    e->profiling = enable != 0;
    return SG_OK;
}

int sg_get_profile(sg_engine* e, double* sweep_ms, uint64_t* sweep_launches, double* gather_ms,
                   uint64_t* gather_launches) {
    SG_REQUIRE(e, "sg_get_profile: NULL engine");
    DeviceGuard g(e->device);
    double ms[2];
    unsigned long long cnt[2];
    e->timer.collect(ms, cnt);
    if (sweep_ms) *sweep_ms = ms[0];
    if (sweep_launches) *sweep_launches = cnt[0];
    if (gather_ms) *gather_ms = ms[1];
    if (gather_launches) *gather_launches = cnt[1];
    return SG_OK;
}

/* development aid: clocks per tcgen05.mma for an operand layout variant (sg_sweep_tc.cu) */
int sg_debug_mma_bench(sg_engine* e, int variant, int n_dim, int iters, long long* host_out2) {
    if (!e || !host_out2) return SG_ERR_INVALID;
    DeviceGuard g(e->device);
    long long* d = nullptr;
    SG_CUDA(cudaMalloc(&d, 16));
    cudaError_t ce = sg::launch_tc_mma_bench(variant, n_dim, iters, d, 0);
    if (ce == cudaSuccess) ce = cudaMemcpy(host_out2, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (ce != cudaSuccess) return fail(SG_ERR_CUDA, "sg_debug_mma_bench", ce);
    return SG_OK;
}

int sg_tc_cluster_size(sg_engine* e) {
    if (!e || !e->Jp || e->R <= 0 || !sg::sweep_tc_supported(e->n, e->n_tc)) return 0;
    return sg::sweep_tc_cluster_size(e->n_tc, e->R);
}

int sg_tc_side_replicas(sg_engine* e, int n_sweeps, int coupling_planes) {
    if (!e || !e->Jp || e->R <= 0 || n_sweeps <= 0) return 0;
    DeviceGuard g(e->device);
    const int planes = coupling_planes ? coupling_planes : (e->planes_needed ? e->planes_needed : 3);
    return sg::sweep_tc_side_replicas(e->n, e->n_tc, planes, e->R, n_sweeps);
}

int sg_query(sg_engine* e, int32_t* n, int32_t* n_pad, int32_t* n_replicas,
             int32_t* max_replicas_per_block, int32_t* sm_count) {
    SG_REQUIRE(e, "sg_query: NULL engine");
    if (n) *n = e->n;
    if (n_pad) *n_pad = e->n_pad;
    if (n_replicas) *n_replicas = e->R;
    if (max_replicas_per_block)
        *max_replicas_per_block = e->n_pad ? sg::sweep_max_replicas_per_block(e->n_pad) : 0;
    if (sm_count) *sm_count = e->sm_count;
    return SG_OK;
}

uint64_t sg_launch_count(sg_engine* e) { return e ? e->launches : 0; }

/* development aid: device buffer of 256*8 int64 clock stamps written by block 0 */
int sg_debug_set_timeline(sg_engine* e, long long* dev_buf) {
    if (!e) return SG_ERR_INVALID;
    e->dbg = dev_buf;
    return SG_OK;
}

}  // extern "C"
