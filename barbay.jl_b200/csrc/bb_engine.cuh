// Host engine: owns the device-resident shard, orchestrates one ADVI step as
// pass 1 -> reduce -> (NCCL all-reduce) -> shared kernel -> pass 2 (-> hyper kernels),
// and implements the parity entry points.  One Engine = one GPU = one host thread.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "bb_derived.cuh"
#include "bb_kernel_set.cuh"
#include "bb_layout.h"

namespace bb {

#define BB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e__) +       \
                                     " at " __FILE__ ":" + std::to_string(__LINE__));              \
    } while (0)

// ------------------------------------------------------------------ NCCL through dlopen
// The library binds the NCCL that torch already loaded in the process (libnccl.so.2), so
// no link-time dependency exists and single-GPU use needs no NCCL at all.
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    typedef void *Comm;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    void *lib = nullptr;
    static NcclApi &get() {
        static NcclApi api;
        if (!api.lib) {
            api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!api.lib) throw std::runtime_error(std::string("cannot load libnccl.so.2: ") + dlerror());
            auto sym = [&](const char *n) {
                void *p = dlsym(api.lib, n);
                if (!p) throw std::runtime_error(std::string("NCCL symbol missing: ") + n);
                return p;
            };
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
            api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
            api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
            api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        }
        return api;
    }
};
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

template <typename T> struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t count, bool zero = true) {
        release();
        n = count;
        if (count == 0) return;
        BB_CUDA(cudaMalloc(&p, count * sizeof(T)));
        if (zero) BB_CUDA(cudaMemset(p, 0, count * sizeof(T)));
    }
    void ensure(size_t count) { if (n < count) alloc(count); }
    void upload(const std::vector<T> &h) {
        alloc(h.size(), false);
        if (!h.empty()) BB_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
constexpr size_t kMaxSmem = 227 * 1024;   // opt-in dynamic shared memory per CTA on sm_100

class EngineBase {
  public:
    virtual ~EngineBase() {}
    std::string err;
    Layout L;
    virtual void init_params(uint64_t seed) = 0;
    virtual void set_params(const double *mu, const double *omega) = 0;
    virtual void get_params(double *mu, double *omega, bool posterior) = 0;
    virtual void logjoint_grad(const double *x, int K, int eps_is_noise, double *logp, double *grad) = 0;
    virtual void elbo_grad(const double *eps, long long step, double *elbo, double *grad) = 0;
    virtual void get_noise(long long step, double *eps) = 0;
    virtual void set_optimizer(const bb_opt &o) = 0;
    virtual void step(int n, double *trace) = 0;
    virtual void step_with_noise(const double *eps) = 0;
    virtual long long state_size() const = 0;
    virtual void get_state(double *s) = 0;
    virtual void set_state(const double *s) = 0;
    virtual void set_stream(void *s) = 0;
    virtual void sync() = 0;
    virtual void time_steps(int n, float *ms_total, float *ms_pass1, float *ms_pass2) = 0;
    virtual void comm_init(const char id[128]) = 0;
    // persistent step kernel: cycles of CTA 0 in {column phase, arrive -> sums, sums -> context} and the number of
    // in-kernel tails, summed since the last call (then reset); out[4] = SM clock in kHz
    virtual void persist_stats(double out[16]) = 0;
    // out = {packed step kernel in use, steps per persistent launch (0: off), peer-memory exchange on, NCCL communicator present}
    virtual void data_plane(int32_t out[8]) = 0;
    // one-process-per-GPU wiring without NCCL: every rank exports the CUDA IPC handle of its exchange buffer, the
    // caller gathers them with whatever transport it has (torch.distributed, MPI, a file) and hands all of them back
    virtual void peer_handle(char out[64]) = 0;
    virtual void peer_attach(const char *handles, int n) = 0;
    // derived bc_fitness rows of the hierarchical models (bb_derived.cuh): median and sd of n draws of
    // theta + exp(log-tau) theta-tilde per log-tau row, [E M R] each; rows of other shards stay 0
    virtual void derived_fitness(int n, uint64_t seed, double *median, double *sd) = 0;
    int home_device = -1;      // CUDA ordinal every call of a single-GPU handle runs on (-1: the handle sets devices itself)
    long long launches = 0;
    long long step_count = 0;
    double alg_bytes = 0.0;
};

EngineBase *make_engine_f32(const bb_desc &d);
EngineBase *make_engine_f64(const bb_desc &d);
EngineBase *make_multi_engine_f32(const bb_desc &d);      // bb_multi.cuh: one handle, n_devices GPUs
EngineBase *make_multi_engine_f64(const bb_desc &d);

// =====================================================================================
template <typename real> class Engine : public EngineBase {
    using r2 = vec2<real>;

  public:
    explicit Engine(const bb_desc &d);
    ~Engine() override;
    void init_params(uint64_t seed) override;
    void set_params(const double *mu, const double *omega) override;
    void get_params(double *mu, double *omega, bool posterior) override;
    void logjoint_grad(const double *x, int K, int eps_is_noise, double *logp, double *grad) override;
    void elbo_grad(const double *eps, long long step, double *elbo, double *grad) override;
    void get_noise(long long step, double *eps) override;
    void set_optimizer(const bb_opt &o) override;
    void step(int n, double *trace) override;
    void step_with_noise(const double *eps) override;
    long long state_size() const override;
    void get_state(double *s) override;
    void set_state(const double *s) override;
    void set_stream(void *s) override { stream_ = s ? (cudaStream_t)s : own_stream_; }
    void sync() override { BB_CUDA(cudaStreamSynchronize(stream_)); }
    void time_steps(int n, float *ms_total, float *ms_pass1, float *ms_pass2) override;
    void comm_init(const char id[128]) override;
    void persist_stats(double out[16]) override;
    void derived_fitness(int n, uint64_t seed, double *median, double *sd) override;
    void peer_handle(char out[64]) override;
    void peer_attach(const char *handles, int n) override;
    void data_plane(int32_t out[8]) override {
        out[0] = stepk_ok_ ? 1 : 0;
        out[1] = (stepk_ok_ && persist_chunk_ > 1 && !(comm_ && !xchg_on_)) ? persist_chunk_ : 0;
        out[2] = xchg_on_ ? 1 : 0;
        out[3] = comm_ ? 1 : 0;
        out[4] = out[5] = out[6] = out[7] = 0;
        if (stepk_ok_) { out[4] = groups_[0].step_occ; out[5] = groups_[0].step_nbuf; out[6] = groups_[0].step_stage_acc; out[7] = groups_[0].stblocks; }
    }
    // single-process multi-GPU (bb_multi.cuh): exchange buffers of all devices of the handle, peer access enabled
    void *xchg_local() const { return xchg_mem_; }
    void attach_peers(const std::vector<void *> &bufs);
    void finish_elbo_partial(double *out);          // [K + 1] this shard's ELBO partial sums (no collective)
    int device() const { return device_; }
    void set_raw_elbo(bool on) { raw_elbo_ = on; }
    void alloc_exchange() { alloc_xchg(); }

  private:
    struct Group {                 // replicates sharing one T (and one environment list): one launch of each column kernel
        int nt = 0;
        unsigned rep_mask = 0;
        int env_of_t[MAX_NT_DYN] = {};
        SegList p1segs, p2segs;
        int p1blocks = 0, p2blocks = 0;
        KernelSet<real> ks, ks_sup;
        int pv = 0, kchunk = 0;
        size_t p1smem = 0, p2smem = 0, p2smem_elbo = 0;
        int p1nbuf = 2, p2nbuf = 2, p2stage_acc = 1, p2stage_ring = 1;
        size_t fsmem = 0;          // fused step kernel: pass-2 staging + pass-1 accumulators
        int facc_slots = 0;        // 0: the fused kernel is not available for this group
        // packed / persistent step kernel (bb_step_kernel.cuh); stepk == nullptr: not available for this shape
        StepKernelFn<real> stepk = nullptr;
        int step_w = 1, step_acc_rows = 0, step_stage_ring = 0, step_occ = 0, step_nbuf = 2, step_stage_acc = 1;
        size_t step_smem = 0;
        SegList stsegs;
        int stblocks = 0;
        bool step_persist = false; // the in-kernel tail's scratch fits behind the flushed accumulators
        size_t part_off = 0, epart_off = 0;     // block offsets into part_ / epart_
    };
    struct RunMode {
        bool sup = false;          // caller-supplied noise
        bool z_direct = false;
        bool update = false;
        bool want_elbo = false;
        bool dump = false;         // per-sample gradients
        bool gout = false;         // (dELBO/dmu, dELBO/domega)
        uint32_t step = 0;
        bool fuse = false;         // pass 2 also accumulates the next step's pass-1 partials
        bool have_part = false;    // this step's partials were produced by the previous fused kernel
        bool stepk = false;        // the fused kernel is the packed step kernel (partials in xpart_, output space)
        bool have_xpart = false;   // this step's partials sit in xpart_
        int nsteps = 1;            // > 1: persistent step kernel covering this many steps
    };

    void build_groups();
    void size_pass2();
    void assign_blocks(SegList &sl, unsigned rep_mask, int budget, int *nblocks);
    void run_pipeline(const RunMode &m);
    void upload_supplied(const double *x, int K);
    void ensure_supplied(bool dump);
    ColArrays<real> col_arrays(bool ring_base = false) const;
    template <typename T> OptArgsT<T> opt_args(bool update) const {
        OptArgsT<T> o;
        o.kind = opt_.kind; o.update = update ? 1 : 0;
        o.eta = (T)opt_.eta; o.tau = (T)opt_.tau; o.post = (T)opt_.post;
        return o;
    }
    void zero_like_init_acc();
    double finish_elbo(double *logp_k);
    void gather_to_ref(const r2 *lam, const r2 *bc, const r2 *hy, const double2 *sh, double *x, double *y, bool sp);
    template <typename F> void launch_count(F &&f) { f(); ++launches; }

    int device_ = 0, nsm_ = 148;
    cudaStream_t stream_ = nullptr, own_stream_ = nullptr;
    uint64_t seed_ = 0;
    bb_opt opt_{BB_OPT_TRUNCATED_ADAGRAD, 0.1, 1.0, 0.9, 100};
    bool opt_ready_ = false;
    int ring_slot_ = 0;
    bool fused_ok_ = false;        // every launch group has a fused step kernel that fits
    long long part_step_ = -1;     // step whose pass-1 partials already sit in part_ (fused layout) / xpart_, or -1
    int hz_k_ = 0, hz_h_ = 1;      // strides of the hyper-latent draws in zeps_: sample-major [K][H] unless BB_HZ_TRANSPOSE=1 ([H][K])
    long long hy_zeps_step_ = -1;  // step whose hyper-latent draws already sit in zeps_ (made by hyper_update_kernel), or -1
    bool part_is_x_ = false;       // ... in xpart_ (written by the step kernel)
    bool stepk_ok_ = false;        // every launch group (there is one: R == 1) has a step kernel that fits
    int persist_chunk_ = 0;        // steps per persistent launch (0: persistent mode off)
    DBuf<double> xpart_, gpart_, eps_steps_;
    DBuf<StepSync> step_sync_;
    int step_gsize_ = 32;
    void alloc_xchg();
    void check_step_sync();
    bool resum_active() const;
    void maybe_resum();
    std::vector<Group> groups_;
    int p1blocks_total_ = 0, p2blocks_total_ = 0, hyblocks_ = 0;

    // device state
    DBuf<r2> lam_th_, lam_acc_, bc_th_, bc_acc_, hy_th_, hy_acc_, lam_ring_, bc_ring_, hy_ring_;
    DBuf<double2> sh_th_, sh_acc_, sh_ring_, sh_pr_, sh_gout_;
    DBuf<r2> lam_pr_, bc_pr_, hy_pr_;
    DBuf<int> cnt_, map_lam_, map_bc_, map_hy_, map_sh_, hgroup_, csr_off_, csr_mem_;
    DBuf<unsigned> ticket_;    // arrival counter of the merged tail kernel (zero between launches)
    DBuf<float2> trig_;        // Box-Muller direction table of the fp32 noise lattice (bb_device.cuh)
    PhiloxKey philox_key(uint64_t seed) const { PhiloxKey k = make_philox_key(seed); k.trig = trig_.p; return k; }
    DBuf<uint32_t> col_id_;
    DBuf<double> part_, sums_, epart_, hy_epart_, elbo_sh_, elbo_out_, sh_scratch_;
    DBuf<real> ctx_;
    DBuf<real> aw_d_, aw_zs_;  // as-written neutral pairing (ragged replicate model): see shared_body
    DBuf<double> aw_tmp_;
    DBuf<r2> zeps_, hcontrib_, hsum_;
    // parity / gradient outputs
    DBuf<real> sup_lam_, sup_bc_, sup_hy_, dump_lam_, dump_bc_, dump_hy_, dump_hc_;
    DBuf<double> sup_sh_, dump_sh_, hostvec_a_, hostvec_b_;
    DBuf<r2> gout_lam_, gout_bc_, gout_hy_;
    DBuf<double> trace_;
    // comm
    NcclApi::Comm comm_ = nullptr;
    // peer-memory exchange (see bb_aux_kernels.cuh): local buffer + IPC-mapped peer buffers
    bool xchg_on_ = false;
    bool raw_elbo_ = false;        // ELBO / log-density values without their constants (summed by a multi-GPU owner)
    bool peers_local_ = false;     // peer buffers are plain peer-access pointers of this process (not IPC mappings)
    unsigned long long xchg_seq_ = 0;
    void *xchg_mem_ = nullptr;
    std::vector<void *> xchg_peer_mem_;
    DBuf<int> xchg_err_;
    size_t xchg_flag_off_ = 0, xchg_aux_off_ = 0, xchg_aux_flag_off_ = 0;
    unsigned long long aux_seq_ = 0;
    static constexpr int kAuxP = 4096;      // doubles per round of the auxiliary all-reduce
    void peer_allreduce(double *dev, size_t n);
    void setup_peer_exchange();
    void check_peer_exchange();
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> tev_;   // per-step kernel brackets while timing
    int tev_pos_ = -1;               // >= 0: record pass brackets into tev_
};

// ------------------------------------------------------------------ construction
template <typename real> Engine<real>::Engine(const bb_desc &d) {
    build_layout(d, L);
    seed_ = d.seed;
    if (d.device >= 0) BB_CUDA(cudaSetDevice(d.device));
    BB_CUDA(cudaGetDevice(&device_));
    home_device = device_;
    cudaDeviceProp prop;
    BB_CUDA(cudaGetDeviceProperties(&prop, device_));
    if (prop.major != 10)
        throw std::runtime_error("barbay_b200 is built for sm_100a (B200) only; device reports sm_" +
                                 std::to_string(prop.major) + std::to_string(prop.minor));
    nsm_ = prop.multiProcessorCount;
    BB_CUDA(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    stream_ = own_stream_;
    for (auto &e : ev_) BB_CUDA(cudaEventCreate(&e));

    {
        // directions sqrt(2 ln 2) (cos, sin)(2 pi (a + 0.5) / TRIG_N), a < TRIG_N (the full turn), rounded once to fp32
        // (the kernel's radius is sqrt(-log2 u); the factor turns it into sqrt(-2 ln u))
        std::vector<float2> t(TRIG_N);
        const double scale = std::sqrt(2.0 * std::log(2.0));
        for (int a = 0; a < TRIG_N; ++a) {
            const double ang = 6.283185307179586476925286766559 * ((double)a + 0.5) / (double)TRIG_N;
            t[a] = make_float2((float)(scale * std::cos(ang)), (float)(scale * std::sin(ang)));
        }
        trig_.upload(t);
        ticket_.alloc(1);
    }
    const size_t nlam = (size_t)L.tmax * L.cpad, nbc = (size_t)L.nj * L.cpad;
    lam_th_.alloc(nlam); lam_acc_.alloc(nlam);
    bc_th_.alloc(nbc); bc_acc_.alloc(nbc);
    cnt_.upload(L.cnt);
    map_lam_.upload(L.map_lam); map_bc_.upload(L.map_bc); map_sh_.upload(L.map_sh);
    if (!L.perm_identity) col_id_.upload(L.col_id);
    sh_th_.alloc(2 * L.nst); sh_acc_.alloc(2 * L.nst); sh_gout_.alloc(2 * L.nst);
    {
        std::vector<double2> p(2 * L.nst);
        for (int i = 0; i < 2 * L.nst; ++i) p[i] = make_double2(L.pr_sh[2 * i], L.pr_sh[2 * i + 1]);
        sh_pr_.upload(p);
    }
    auto to_r2 = [](const std::vector<double> &v) {
        std::vector<r2> o(v.size() / 2);
        for (size_t i = 0; i < o.size(); ++i) o[i] = r2{(real)v[2 * i], (real)v[2 * i + 1]};
        return o;
    };
    if (L.lam_pr_matrix) lam_pr_.upload(to_r2(L.pr_lam));
    if (L.bc_pr_matrix) bc_pr_.upload(to_r2(L.pr_bc));
    if (L.hier) {
        hy_th_.alloc(L.H); hy_acc_.alloc(L.H);
        hy_pr_.upload(to_r2(L.pr_hy));
        map_hy_.upload(L.map_hy);
        hgroup_.upload(L.hgroup);
        csr_off_.upload(L.csr_off); csr_mem_.upload(L.csr_mem);
        zeps_.alloc((size_t)L.K * std::max(L.H, 1));
        if (getenv("BB_HZ_TRANSPOSE")) { hz_k_ = 1; hz_h_ = L.K; } else { hz_k_ = std::max(L.H, 1); hz_h_ = 1; }
        hcontrib_.alloc((size_t)L.E * L.cpad);
        hyblocks_ = cdiv(std::max(L.H, 1), BLOCK);
        hy_epart_.alloc((size_t)hyblocks_ * (L.K + 1));
    }
    sums_.alloc((size_t)L.R * L.K * NQ * L.tmax);
    ctx_.alloc((size_t)L.R * L.K * 3 * L.tmax);
    sh_scratch_.alloc((size_t)3 * L.K * 2 * L.nst + (size_t)2 * L.R * L.K * L.tmax + 2 * L.nst);
    elbo_sh_.alloc(L.K + 1);
    elbo_out_.alloc(L.K + 1);
    if (L.as_written) {            // buffers of the as-written neutral pairing (bb_aux_kernels.cuh, shared_body)
        aw_d_.alloc((size_t)L.R * L.K * std::max(L.N, 1) * (L.tmax - 1));
        aw_zs_.alloc((size_t)L.R * L.K * L.tmax);
        aw_tmp_.alloc((size_t)3 * L.R * L.K * L.tmax);
    }
    build_groups();

    // algorithmic bytes per step of this shard: theta + accumulators read and written
    // once, int32 counts read once; matrix priors read once.
    const double w = sizeof(real);
    double dloc = 0;   // latents owned by the shard
    for (int m : L.map_lam) dloc += m >= 0;
    for (int m : L.map_bc) dloc += m >= 0;
    dloc += L.H + (L.rank == 0 ? 2.0 * L.nst : 0.0);
    double ncnt = 0;
    for (const HostSeg &s : L.segs) ncnt += (double)s.ncol * s.nt;
    alg_bytes = 8.0 * w * dloc + 4.0 * ncnt;
    if (L.lam_pr_matrix) alg_bytes += 2.0 * w * ncnt;
    bb_opt def{BB_OPT_TRUNCATED_ADAGRAD, 0.1, 1.0, 0.9, 100};
    opt_ = def;
}

template <typename real> Engine<real>::~Engine() {
    for (size_t r = 0; r < xchg_peer_mem_.size() && !peers_local_; ++r)
        if (xchg_peer_mem_[r] && (int)r != L.rank) cudaIpcCloseMemHandle(xchg_peer_mem_[r]);
    if (xchg_mem_) cudaFree(xchg_mem_);
    if (comm_) NcclApi::get().CommDestroy(comm_);
    for (auto &e : ev_) if (e) cudaEventDestroy(e);
    for (auto &e : tev_) cudaEventDestroy(e);
    if (own_stream_) cudaStreamDestroy(own_stream_);
}

template <typename real> void Engine<real>::assign_blocks(SegList &sl, unsigned rep_mask, int budget, int *nblocks) {
    // proportional split of the persistent grid among the segments of this launch
    long long tiles_total = 0;
    std::vector<int> tiles;
    for (const HostSeg &s : L.segs)
        if ((rep_mask >> s.rep) & 1u) { tiles.push_back(cdiv(s.ncol, BLOCK)); tiles_total += tiles.back(); }
    sl.nseg = 0;
    int blk = 0, idx = 0;
    for (const HostSeg &s : L.segs) {
        if (!((rep_mask >> s.rep) & 1u)) continue;
        Seg g;
        g.col0 = s.col0; g.ncol = s.ncol; g.rep = s.rep; g.nt = s.nt; g.neutral = s.neutral;
        g.colid0 = s.colid0; g.sh0 = s.sh0;
        long long want = tiles_total ? (long long)budget * tiles[idx] / tiles_total : 1;
        int nb = (int)std::max<long long>(1, std::min<long long>(want, tiles[idx]));
        g.blk0 = blk; g.blk1 = blk + nb;
        blk += nb;
        sl.seg[sl.nseg++] = g;
        ++idx;
    }
    *nblocks = blk;
}

template <typename real> void Engine<real>::build_groups() {
    // one launch group per distinct (T_r, environment list of replicate r): the kernels take one env_of_t per launch
    std::vector<int> lead;         // first replicate of every group
    auto same = [&](int a, int b) { return L.nt[a] == L.nt[b] && L.env_of_rt[a] == L.env_of_rt[b]; };
    for (int r = 0; r < L.R; ++r) {
        bool found = false;
        for (int l : lead) found = found || same(l, r);
        if (!found) lead.push_back(r);
    }
    size_t part_off = 0;
    for (int l : lead) {
        Group g;
        const int nt = L.nt[l];
        g.nt = nt;
        for (int r = 0; r < L.R; ++r) if (same(l, r)) g.rep_mask |= 1u << r;
        for (int t = 0; t < MAX_NT_DYN; ++t) g.env_of_t[t] = L.env_of_rt[l][t];
        if (!lookup_kernels<real>(nt, L.E, L.hier, false, &g.ks) ||
            !lookup_kernels<real>(nt, L.E, L.hier, true, &g.ks_sup))
            throw std::runtime_error("no kernel instantiation for this shape");
        g.pv = nt + 2 * (nt - 1);
        // shared-memory accumulator budget of pass 1: sized for the mutant population (2 nt slots per
        // sample when E == 1), capped so >= 4 CTAs fit per SM; blocks sweep K in chunks when it is short
        const int pv_m = L.E == 1 ? 2 * nt : g.pv;
        const size_t slot_b = (size_t)BLOCK * sizeof(real);
        int slots = L.K * pv_m;
        // BB_P1_ACC_KB (tuning knob): shared-memory budget of the pass-1 accumulators
        const char *acc_kb_env = getenv("BB_P1_ACC_KB");
        const int acc_kb = acc_kb_env ? std::max(4, atoi(acc_kb_env)) : 44;
        const int max_slots = (int)((acc_kb * 1024) / slot_b);
        if (slots > max_slots) {
            const int sweeps = (slots + max_slots - 1) / max_slots;
            slots = ((L.K + sweeps - 1) / sweeps) * pv_m;
        }
        slots = std::max(slots, (g.pv + 1) & ~1);   // at least one sample of the widest population (pair layout: even)
        g.kchunk = slots;
        // + two cp.async staging buffers (see bb_kernels.cuh)
        const size_t th_b = (size_t)(L.tmax + L.nj) * BLOCK * sizeof(r2);
        const size_t p1acc = ((size_t)g.kchunk * slot_b + 15) / 16 * 16;
        const size_t trig_b = TrigTab<real>::BYTES;      // fp32: Box-Muller direction table at the head of smem
        g.p1nbuf = trig_b + p1acc + 2 * th_b <= kMaxSmem ? 2 : 1;
        g.p1smem = trig_b + p1acc + g.p1nbuf * th_b;
        if (g.p1smem > kMaxSmem) throw std::runtime_error("T x E too large for the column kernels' shared memory");
        BB_CUDA(cudaFuncSetAttribute((const void *)g.ks.pass1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.p1smem));
        BB_CUDA(cudaFuncSetAttribute((const void *)g.ks_sup.pass1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.p1smem));
        int occ1 = 1;
        BB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, g.ks.pass1, BLOCK, g.p1smem));
        assign_blocks(g.p1segs, g.rep_mask, nsm_ * std::max(occ1, 1), &g.p1blocks);
        g.part_off = part_off;
        part_off += g.p1blocks;
        groups_.push_back(g);
    }
    p1blocks_total_ = (int)part_off;
    size_pass2();
}

// pass-2 shared memory, occupancy and grid: depends on the optimiser (ring staging), so it is redone
// by set_optimizer
template <typename real> void Engine<real>::size_pass2() {
    const size_t th_b = (size_t)(L.tmax + L.nj) * BLOCK * sizeof(r2);
    // fixed head of the pass-2 shared memory: fp32 Box-Muller direction table + the per-sample context
    constexpr int kVW = 16 / (int)sizeof(real);           // context rows are padded to 16 bytes (bb_kernels.cuh)
    const size_t ctx_b = TrigTab<real>::BYTES + (size_t)L.K * (((3 * L.tmax + kVW - 1) / kVW) * kVW) * sizeof(real);
    const int npr = (L.lam_pr_matrix || L.bc_pr_matrix) ? 1 : 0;
    const int nrg = opt_.kind == BB_OPT_TRUNCATED_ADAGRAD ? 1 : 0;
    size_t epart_off = 0;
    for (Group &g : groups_) {
        // staging level by shared-memory budget: (2 buffers, everything) -> (1 buffer, everything) ->
        // (1 buffer, theta + counts only; accumulators / priors / ring read from global)
        const size_t cn_b = (size_t)L.tmax * BLOCK * sizeof(int);
        const size_t sel_b = (size_t)(L.K + 1) * BLOCK * sizeof(double);
        // prefetched set (theta, priors, counts: one or two buffers) + epilogue set (accumulators, ring: one).
        // TruncatedADAGrad: the evicted ring slot is staged too unless that costs the fused step kernel its third
        // resident CTA per SM (cfg2, K = 8) -- then the epilogue reads the slot straight from global memory.
        auto plan = [&](int stage_ring) {
            const size_t pre_b = (1 + npr) * th_b + cn_b, epi_b = (1 + stage_ring) * th_b;
            g.p2nbuf = 2; g.p2stage_acc = 1; g.p2stage_ring = stage_ring;
            size_t stage_b = 2 * pre_b + epi_b;
            if (ctx_b + sel_b + stage_b > kMaxSmem) { g.p2nbuf = 1; stage_b = pre_b + epi_b; }
            if (ctx_b + sel_b + stage_b > kMaxSmem) { g.p2stage_acc = 0; stage_b = th_b + cn_b; }
            if (ctx_b + sel_b + stage_b > kMaxSmem)
                throw std::runtime_error("T x E too large for the column kernels' shared memory");
            g.p2smem = ctx_b + stage_b;                                       // ELBO = false kernels
            g.p2smem_elbo = g.p2smem + sel_b;
            // fused step kernel (non-hierarchical, all K samples in one sweep of the mutant accumulators)
            g.facc_slots = 0;
            if (g.ks.pass2_fused && !getenv("BB_NO_FUSE")) {
                const int pv_m = L.E == 1 ? 2 * g.nt : g.pv;
                const int slots = std::max(L.K * pv_m, (g.pv + 1) & ~1);
                const size_t fs = g.p2smem + (size_t)slots * BLOCK * sizeof(real);
                if (fs <= kMaxSmem && (size_t)slots * BLOCK * sizeof(real) <= 64 * 1024) {
                    BB_CUDA(cudaFuncSetAttribute((const void *)g.ks.pass2_fused,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs));
                    int occf = 0;
                    BB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occf, g.ks.pass2_fused, BLOCK, fs));
                    // measured (B200, cfg2): the fused kernel wins at >= 3 resident CTAs per SM; with 2 the
                    // two-kernel step is faster
                    if (occf >= 3) { g.facc_slots = slots; g.fsmem = fs; }
                }
            }
        };
        plan(nrg);
        if (nrg && !g.facc_slots && g.ks.pass2_fused && !L.hier) {
            plan(0);
            if (!g.facc_slots) plan(nrg);          // no fused kernel either way: keep the ring staged
        }
        // ---- packed step kernel (bb_step_kernel.cuh): non-hierarchical models, compile-time T and E
        g.stepk = nullptr;
        if (!L.hier && L.R == 1 && !getenv("BB_NO_STEPK")) {
            const int W = (L.K % 2 == 0) ? 2 : 1;
            StepKernelFn<real> fn = W == 2 ? g.ks.step_w2 : g.ks.step_w1;
            // odd K (the reference's default K = 1) on one GPU: the round-1 fused kernel keeps 4 CTAs per SM and is the
            // faster one there (measured: 59.5 vs 67.8 us at cfg2); the unpacked step kernel serves multi-GPU shards
            if (W == 1 && L.world == 1 && !getenv("BB_STEPK_ODD")) fn = nullptr;
            if (fn) {
                auto r128 = [](size_t b) { return (b + 127) / 128 * 128; };
                const int NJ = 2 * L.E, ROWS = g.nt + NJ, npack = L.K / W;
                const size_t thb = (size_t)ROWS * BLOCK * sizeof(r2), cnb = (size_t)g.nt * BLOCK * sizeof(int);
                const int pvs_m = L.E == 1 ? 2 * g.nt : 3 * g.nt - 2, pvs_n = 3 * g.nt - 2;
                const int rows = std::max(npack * ((pvs_m + 1) / 2), (pvs_n + 1) / 2);
                const size_t spb = (size_t)2 * W * sizeof(real);        // one SlotPair
                // mbarriers | packed context | ... | behind the accumulators: (theta, acc) and sigma of the population latents
                const size_t head = 128 + r128((size_t)npack * 3 * g.nt * W * sizeof(real)) +
                                    (size_t)4 * (g.nt - 1) * sizeof(double2) + (size_t)2 * (g.nt - 1) * sizeof(real);
                auto total = [&](int nbuf, int stage_acc, int stage_ring) {
                    return head + (size_t)nbuf * ((1 + npr) * thb + cnb) +
                           (stage_acc ? (size_t)(1 + stage_ring) * thb : 0) + (size_t)rows * BLOCK * spb;
                };
                const int min_occ = getenv("BB_STEPK_MIN_OCC") ? atoi(getenv("BB_STEPK_MIN_OCC")) : 3;
                // staging levels, richest first: (double buffer, accumulators [+ ring] staged) ... (single buffer, none).
                // The level with the most resident CTAs per SM wins (the step is latency-bound: warps, not prefetch depth,
                // are what it lacks); ties go to the richer level.  BB_STEPK_STAGE=<nbuf><acc><ring> forces one.
                int best_occ = 0;
                const char *force = getenv("BB_STEPK_STAGE");
                for (int lvl = 0; lvl < 6; ++lvl) {
                    static const int cfgs[6][3] = {{2, 1, 1}, {2, 1, 0}, {1, 1, 1}, {1, 1, 0}, {2, 0, 0}, {1, 0, 0}};
                    const int nbuf = cfgs[lvl][0], sacc = cfgs[lvl][1], sring = cfgs[lvl][2];
                    if (sring && !nrg) continue;
                    if (force && (force[0] - '0' != nbuf || force[1] - '0' != sacc || force[2] - '0' != sring)) continue;
                    const size_t sm = total(nbuf, sacc, sring);
                    if (sm > kMaxSmem) continue;
                    BB_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                    int occ = 0;
                    BB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, BLOCK, sm));
                    if (occ >= min_occ && occ > best_occ) {
                        best_occ = occ;
                        g.stepk = fn; g.step_w = W; g.step_acc_rows = rows; g.step_stage_ring = sring;
                        g.step_nbuf = nbuf; g.step_stage_acc = sacc;
                        g.step_occ = occ; g.step_smem = sm;
                    }
                }
                if (g.stepk)
                    BB_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.step_smem));
                if (g.stepk) {
                    // BB_STEPK_CTAS (tuning): use fewer resident CTAs per SM than fit
                    const int use_occ = getenv("BB_STEPK_CTAS") ? std::max(1, std::min(g.step_occ, atoi(getenv("BB_STEPK_CTAS")))) : g.step_occ;
                    assign_blocks(g.stsegs, g.rep_mask, nsm_ * use_occ, &g.stblocks);
                    // scratch of the in-kernel tail (totals + shared_body's working arrays) aliases the accumulators
                    const size_t tail_d = sums_.n + (size_t)L.K * (2 * g.nt + 7 * (g.nt - 1)) + 16;   // totals + working arrays
                    g.step_persist = tail_d * sizeof(double) <= (size_t)rows * BLOCK * spb && (int)sums_.n <= 4096 &&
                                     L.K * 2 * (g.nt - 1) <= BLOCK;
                }
            }
        }
        BB_CUDA(cudaFuncSetAttribute((const void *)g.ks.pass2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)g.p2smem));
        for (auto *fn : {(const void *)g.ks.pass2_elbo, (const void *)g.ks_sup.pass2_elbo})
            BB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.p2smem_elbo));
        int occ2 = 1;
        if (g.facc_slots)
            BB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, g.ks.pass2_fused, BLOCK, g.fsmem));
        else
            BB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, g.ks.pass2, BLOCK, g.p2smem));
        assign_blocks(g.p2segs, g.rep_mask, nsm_ * std::max(occ2, 1), &g.p2blocks);
        g.epart_off = epart_off;
        epart_off += g.p2blocks;
    }
    p2blocks_total_ = (int)epart_off;
    epart_.alloc((size_t)p2blocks_total_ * (L.K + 1));
    part_.alloc((size_t)std::max(p1blocks_total_, p2blocks_total_) * L.K * (3 * L.tmax));
    fused_ok_ = !L.hier;
    for (Group &g : groups_) fused_ok_ = fused_ok_ && g.facc_slots > 0;
    stepk_ok_ = !L.hier && groups_.size() == 1 && groups_[0].stepk != nullptr;
    persist_chunk_ = 0;
    if (stepk_ok_) {
        Group &g = groups_[0];
        xpart_.alloc((size_t)g.stblocks * sums_.n);

        int coop = 0;
        BB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device_));
        // default: on (256 steps per launch), on one GPU as on the shards of a multi-GPU run: 147 us per step at cfg2
        // against 150-154 us for the launch pair (profiles/r2_smallshard.jsonl, r2_bench_n1.json), and every GPU count
        // then runs the same code.  Its in-kernel shared-latent phases use the kernel's own precision and operation
        // order, so a run split into several bb_step calls (or resumed from a checkpoint) agrees with the uninterrupted
        // one to rounding, not bitwise; BB_PERSIST=0 selects the launch pair, which is bitwise split-invariant.
        const char *pe = getenv("BB_PERSIST");
        int want = pe ? atoi(pe) : 256;
        if (!coop || !g.step_persist) want = 0;
        if (L.world > 1 && !xchg_on_) want = 0;
        if (want > 1) {
            persist_chunk_ = want;
            step_gsize_ = getenv("BB_STEPK_GSIZE") ? std::max(8, atoi(getenv("BB_STEPK_GSIZE"))) : 32;
            while (cdiv(g.stblocks, step_gsize_) > 64) step_gsize_ *= 2;
            gpart_.alloc((size_t)cdiv(g.stblocks, step_gsize_) * sums_.n);
            if (!step_sync_.p) step_sync_.alloc(1);
            eps_steps_.alloc((size_t)persist_chunk_ * L.K * 2 * L.nst);
            alloc_xchg();
            if (xchg_peer_mem_.empty() && L.world == 1) xchg_peer_mem_.assign(1, xchg_mem_);
        }
    }
    part_step_ = -1;
    hy_zeps_step_ = -1;
}

template <typename real> ColArrays<real> Engine<real>::col_arrays(bool ring_base) const {
    ColArrays<real> C;
    C.cpad = L.cpad; C.tmax = L.tmax; C.nj = L.nj;
    C.lam_th = lam_th_.p; C.lam_acc = lam_acc_.p; C.cnt = cnt_.p;
    C.bc_th = bc_th_.p; C.bc_acc = bc_acc_.p;
    C.lam_pr = lam_pr_.p; C.bc_pr = bc_pr_.p;
    C.lam_pr_s = r2{(real)L.pr_lam_s[0], (real)L.pr_lam_s[1]};
    for (int k = 0; k < 3; ++k) C.bc_pr_s[k] = r2{(real)L.pr_bc_s[k][0], (real)L.pr_bc_s[k][1]};
    C.hgroup = hgroup_.p; C.col_id = col_id_.p;
    // the rings are [n][...]: hand the kernels this step's slot
    const int slot = ring_base ? 0 : ring_slot_;
    C.lam_ring = lam_ring_.p ? lam_ring_.p + (size_t)slot * L.tmax * L.cpad : nullptr;
    C.bc_ring = bc_ring_.p ? bc_ring_.p + (size_t)slot * L.nj * L.cpad : nullptr;
    return C;
}

// ------------------------------------------------------------------ parameters
template <typename real> void Engine<real>::init_params(uint64_t seed) {
    part_step_ = -1;
    hy_zeps_step_ = -1;
    const PhiloxKey ikey = philox_key(seed);
    auto run = [&](r2 *dst, const int *map, size_t n) {
        if (!n) return;
        init_params_kernel<real><<<cdiv(n, 256), 256, 0, stream_>>>(dst, map, (long long)n, L.D, ikey);
        ++launches;
    };
    run(lam_th_.p, map_lam_.p, lam_th_.n);
    run(bc_th_.p, map_bc_.p, bc_th_.n);
    if (L.hier) run(hy_th_.p, map_hy_.p, hy_th_.n);
    // shared latents are replicated on every rank: generate from the identity map
    std::vector<double2> sh(2 * L.nst);
    {
        DBuf<int> idm;
        std::vector<int> id(2 * L.nst);
        for (int i = 0; i < 2 * L.nst; ++i) id[i] = i;
        idm.upload(id);
        DBuf<double2> tmp;
        tmp.alloc(2 * L.nst);
        init_params_kernel<double><<<cdiv(2 * L.nst, 128), 128, 0, stream_>>>(tmp.p, idm.p, 2 * L.nst, L.D, ikey);
        ++launches;
        BB_CUDA(cudaMemcpyAsync(sh_th_.p, tmp.p, sizeof(double2) * 2 * L.nst, cudaMemcpyDeviceToDevice, stream_));
        BB_CUDA(cudaStreamSynchronize(stream_));
    }
    BB_CUDA(cudaGetLastError());
}

template <typename real> void Engine<real>::set_params(const double *mu, const double *omega) {
    part_step_ = -1;
    hy_zeps_step_ = -1;
    hostvec_a_.ensure(L.D); hostvec_b_.ensure(L.D);
    BB_CUDA(cudaMemcpyAsync(hostvec_a_.p, mu, sizeof(double) * L.D, cudaMemcpyHostToDevice, stream_));
    BB_CUDA(cudaMemcpyAsync(hostvec_b_.p, omega, sizeof(double) * L.D, cudaMemcpyHostToDevice, stream_));
    auto run = [&](r2 *dst, const int *map, size_t n) {
        if (!n) return;
        scatter_pairs_kernel<real><<<cdiv(n, 256), 256, 0, stream_>>>(dst, map, (long long)n, hostvec_a_.p,
                                                                      hostvec_b_.p, 0.0, 0.0);
        ++launches;
    };
    run(lam_th_.p, map_lam_.p, lam_th_.n);
    run(bc_th_.p, map_bc_.p, bc_th_.n);
    if (L.hier) run(hy_th_.p, map_hy_.p, hy_th_.n);
    // shared latents: first 2 nst reference entries, on every rank
    std::vector<double2> sh(2 * L.nst);
    for (int i = 0; i < 2 * L.nst; ++i) sh[i] = make_double2(mu[i], omega[i]);
    BB_CUDA(cudaMemcpyAsync(sh_th_.p, sh.data(), sizeof(double2) * sh.size(), cudaMemcpyHostToDevice, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    BB_CUDA(cudaGetLastError());
}

template <typename real>
void Engine<real>::gather_to_ref(const r2 *lam, const r2 *bc, const r2 *hy, const double2 *sh, double *x, double *y,
                                 bool sp) {
    // x, y: device vectors of length D, zero-filled here; entries not owned by this shard stay 0
    if (x) BB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * L.D, stream_));
    if (y) BB_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * L.D, stream_));
    auto run = [&](const r2 *src, const int *map, size_t n) {
        if (!n || !src) return;
        gather_pairs_kernel<real><<<cdiv(n, 256), 256, 0, stream_>>>(src, map, (long long)n, x, y, sp ? 1 : 0);
        ++launches;
    };
    run(lam, map_lam_.p, lam_th_.n);
    run(bc, map_bc_.p, bc_th_.n);
    if (L.hier) run(hy, map_hy_.p, hy_th_.n);
    if (sh && L.nst) {
        gather_pairs_kernel<double><<<cdiv(2 * L.nst, 128), 128, 0, stream_>>>(sh, map_sh_.p, 2 * L.nst, x, y,
                                                                               sp ? 1 : 0);
        ++launches;
    }
}

template <typename real> void Engine<real>::get_params(double *mu, double *omega, bool posterior) {
    check_peer_exchange();
    hostvec_a_.ensure(L.D); hostvec_b_.ensure(L.D);
    gather_to_ref(lam_th_.p, bc_th_.p, hy_th_.p, sh_th_.p, hostvec_a_.p, hostvec_b_.p, posterior);
    BB_CUDA(cudaMemcpyAsync(mu, hostvec_a_.p, sizeof(double) * L.D, cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaMemcpyAsync(omega, hostvec_b_.p, sizeof(double) * L.D, cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    BB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ optimiser
template <typename real> void Engine<real>::set_optimizer(const bb_opt &o) {
    if (o.kind != BB_OPT_TRUNCATED_ADAGRAD && o.kind != BB_OPT_DECAYED_ADAGRAD)
        throw std::runtime_error("opt must be TruncatedADAGrad or DecayedADAGrad");
    if (o.kind == BB_OPT_TRUNCATED_ADAGRAD && o.n < 1) throw std::runtime_error("TruncatedADAGrad needs n >= 1");
    opt_ = o;
    step_count = 0;
    ring_slot_ = 0;
    size_pass2();
    const size_t nlam = lam_th_.n, nbc = bc_th_.n, nhy = hy_th_.n, nsh = sh_th_.n;
    if (o.kind == BB_OPT_TRUNCATED_ADAGRAD) {
        lam_ring_.alloc(nlam * o.n); bc_ring_.alloc(nbc * o.n);
        if (L.hier) hy_ring_.alloc(nhy * o.n);
        sh_ring_.alloc(nsh * ((size_t)o.n + 1));     // n + 1 slots: see SharedArgs (bb_aux_kernels.cuh)
        BB_CUDA(cudaMemsetAsync(lam_acc_.p, 0, sizeof(r2) * nlam, stream_));
        BB_CUDA(cudaMemsetAsync(bc_acc_.p, 0, sizeof(r2) * nbc, stream_));
        if (nhy) BB_CUDA(cudaMemsetAsync(hy_acc_.p, 0, sizeof(r2) * nhy, stream_));
        BB_CUDA(cudaMemsetAsync(sh_acc_.p, 0, sizeof(double2) * nsh, stream_));
    } else {
        lam_ring_.release(); bc_ring_.release(); hy_ring_.release(); sh_ring_.release();
        const r2 e{(real)1e-8, (real)1e-8};                 // acc = fill(1e-8) (AdvancedVI optimisers.jl)
        fill_kernel<r2><<<cdiv(nlam, 256), 256, 0, stream_>>>(lam_acc_.p, (long long)nlam, e);
        fill_kernel<r2><<<cdiv(nbc, 256), 256, 0, stream_>>>(bc_acc_.p, (long long)nbc, e);
        if (nhy) fill_kernel<r2><<<cdiv(nhy, 256), 256, 0, stream_>>>(hy_acc_.p, (long long)nhy, e);
        fill_kernel<double2><<<cdiv(nsh, 128), 128, 0, stream_>>>(sh_acc_.p, (long long)nsh, make_double2(1e-8, 1e-8));
        launches += 3 + (nhy ? 1 : 0);
    }
    BB_CUDA(cudaStreamSynchronize(stream_));
    BB_CUDA(cudaGetLastError());
    opt_ready_ = true;
    // TruncatedADAGrad (running-sum form) adds an evicted-slot read and a new-slot write per component
    const double w = sizeof(real);
    double dloc = 0;
    for (int m : L.map_lam) dloc += m >= 0;
    for (int m : L.map_bc) dloc += m >= 0;
    dloc += L.H + (L.rank == 0 ? 2.0 * L.nst : 0.0);
    double ncnt = 0;
    for (const HostSeg &s : L.segs) ncnt += (double)s.ncol * s.nt;
    alg_bytes = (o.kind == BB_OPT_TRUNCATED_ADAGRAD ? 12.0 : 8.0) * w * dloc + 4.0 * ncnt;
    if (L.lam_pr_matrix) alg_bytes += 2.0 * w * ncnt;
}

// ------------------------------------------------------------------ supplied noise plumbing
template <typename real> void Engine<real>::ensure_supplied(bool dump) {
    const size_t K = L.K;
    sup_lam_.ensure(K * lam_th_.n); sup_bc_.ensure(K * std::max<size_t>(bc_th_.n, 1));
    sup_sh_.ensure(K * 2 * L.nst);
    if (L.hier) sup_hy_.ensure(K * std::max<size_t>(hy_th_.n, 1));
    if (dump) {
        dump_lam_.ensure(K * lam_th_.n); dump_bc_.ensure(K * std::max<size_t>(bc_th_.n, 1));
        dump_sh_.ensure(K * 2 * L.nst);
        if (L.hier) { dump_hy_.ensure(K * std::max<size_t>(hy_th_.n, 1)); dump_hc_.ensure(K * (size_t)L.E * L.cpad); }
    }
}

template <typename real> void Engine<real>::upload_supplied(const double *x, int K) {
    // x: host [K][D] reference order -> device layout rows
    hostvec_a_.ensure((size_t)K * L.D);
    BB_CUDA(cudaMemcpyAsync(hostvec_a_.p, x, sizeof(double) * K * L.D, cudaMemcpyHostToDevice, stream_));
    auto run = [&](real *dst, const int *map, size_t n) {
        if (!n) return;
        scatter_rows_kernel<real><<<cdiv(n, 256), 256, 0, stream_>>>(dst, map, (long long)n, K, hostvec_a_.p, L.D);
        ++launches;
    };
    run(sup_lam_.p, map_lam_.p, lam_th_.n);
    run(sup_bc_.p, map_bc_.p, bc_th_.n);
    if (L.hier) run(sup_hy_.p, map_hy_.p, hy_th_.n);
    // shared: the first 2 nst entries of each row (every rank)
    BB_CUDA(cudaMemcpy2DAsync(sup_sh_.p, sizeof(double) * 2 * L.nst, hostvec_a_.p, sizeof(double) * L.D,
                              sizeof(double) * 2 * L.nst, K, cudaMemcpyDeviceToDevice, stream_));
}

// ------------------------------------------------------------------ the step pipeline
template <typename real> void Engine<real>::run_pipeline(const RunMode &m) {
    if (!m.fuse) part_step_ = -1;     // anything but a fused step invalidates the pipelined partial sums
    const PhiloxKey pkey = philox_key(seed_);
    const ColArrays<real> C = col_arrays();
    SupArgs<real> sup{};
    if (m.sup) {
        sup.eps_lam = sup_lam_.p; sup.eps_bc = sup_bc_.p; sup.z_direct = m.z_direct ? 1 : 0;
        if (m.dump) { sup.dump_lam = dump_lam_.p; sup.dump_bc = dump_bc_.p; sup.dump_hcontrib = L.hier ? dump_hc_.p : nullptr; }
    }
    HyperArgs<real> ha{};
    if (L.hier) {
        ha.H = L.H; ha.K = L.K; ha.gid0 = L.hy_gid0;
        ha.hy_th = hy_th_.p; ha.hy_acc = hy_acc_.p;
        ha.hy_ring = hy_ring_.p ? hy_ring_.p + (size_t)ring_slot_ * L.H : nullptr; ha.hy_pr = hy_pr_.p;
        ha.zeps = zeps_.p; ha.hz_k = hz_k_; ha.hz_h = hz_h_; ha.key = pkey; ha.step = m.step;
        ha.eps_hy = m.sup ? sup_hy_.p : nullptr; ha.z_direct = m.z_direct ? 1 : 0;
        ha.csr_off = csr_off_.p; ha.csr_mem = csr_mem_.p; ha.hcontrib = hcontrib_.p;
        ha.dump_hcontrib = m.dump ? dump_hc_.p : nullptr; ha.dump_stride = (long long)L.E * L.cpad;
        ha.dump = m.dump ? dump_hy_.p : nullptr;
        ha.gout = m.gout ? gout_hy_.p : nullptr;
        ha.epart = m.want_elbo ? hy_epart_.p : nullptr;
        ha.opt = opt_args<real>(m.update);
        // the draws of this step are already there when the previous step's hyper_update_kernel made them
        const bool have_zeps = !m.sup && hy_zeps_step_ == (long long)m.step;
        ha.prep_next = (m.update && !m.sup && !m.z_direct && !getenv("BB_NO_HYPER_MERGE")) ? 1 : 0;
        if (L.H > 0 && !have_zeps) {
            hyper_prep_kernel<real><<<hyblocks_, BLOCK, 0, stream_>>>(ha);
            ++launches;
        }
        hy_zeps_step_ = ha.prep_next ? (long long)m.step + 1 : -1;
    }
    // ---- pass 1
    if (tev_pos_ >= 0) BB_CUDA(cudaEventRecord(tev_[tev_pos_++], stream_));
    // one launch group and no NCCL collective in between: reduce + peer post + shared latents run as ONE
    // kernel with its working arrays in shared memory (tail_kernel); otherwise the separate kernels
    const int nwarps = L.R * L.K * NQ * L.tmax;
    const size_t tail_smem = (sums_.n + sh_scratch_.n) * sizeof(double);
    if (L.world > 1 && !comm_ && !xchg_on_)
        throw std::runtime_error("this handle is shard " + std::to_string(L.rank) + " of " + std::to_string(L.world) +
                                 ": call bb_comm_init before any evaluation (the step's sums must be combined)");
    const bool use_tail = groups_.size() == 1 && !(comm_ && !xchg_on_) && tail_smem <= 40 * 1024 && !getenv("BB_NO_TAIL");
    ReduceArgs ra{};
    for (Group &g : groups_) {
        // partial sums of this step: from the previous fused kernel (its grid / segment split), or pass 1 now
        double *gpart = part_.p + (size_t)(m.have_part ? g.epart_off : g.part_off) * L.K * (3 * L.tmax);
        if (!m.have_part && !m.have_xpart) {
            P1Args<real> a{};
            a.segs = g.p1segs; a.cols = C; a.K = L.K; a.acc_slots = g.kchunk; a.ne = L.E;
            for (int t = 0; t < MAX_NT_DYN; ++t) a.env_of_t[t] = g.env_of_t[t];
            a.key = pkey; a.step = m.step;
            a.hy_zeps = zeps_.p; a.H = L.H; a.hz_k = hz_k_; a.hz_h = hz_h_;
            a.part = gpart;
            a.pv = g.pv; a.sup = sup; a.nbuf = g.p1nbuf;
            (m.sup ? g.ks_sup.pass1 : g.ks.pass1)<<<g.p1blocks, BLOCK, g.p1smem, stream_>>>(a);
            ++launches;
        }
        if (tev_pos_ >= 0 && &g == &groups_.back()) BB_CUDA(cudaEventRecord(tev_[tev_pos_++], stream_));
        ra = ReduceArgs{};
        ra.segs = m.have_part ? g.p2segs : g.p1segs; ra.K = L.K; ra.tmax = L.tmax; ra.pv = g.pv; ra.nt = g.nt;
        ra.w_single = L.E == 1 ? 1 : 0; ra.rep_mask = g.rep_mask; ra.nblk = m.have_part ? g.p2blocks : g.p1blocks;
        ra.part = gpart; ra.sums = sums_.p;
        if (m.have_xpart) { ra.xpart = xpart_.p; ra.nblk = g.stblocks; }
        if (!use_tail) {
            reduce_kernel<<<cdiv((long long)nwarps * 32, 128), 128, 0, stream_>>>(ra, L.R);
            ++launches;
        }
    }
    if (L.as_written) {
        // as-written neutral pairing (ragged replicate model): the neutral columns' log-ratio differences of this step
        AwArgs<real> aw{};
        for (const HostSeg &s : L.segs) {
            if (!s.neutral) continue;
            Seg g{};
            g.col0 = s.col0; g.ncol = s.ncol; g.rep = s.rep; g.nt = s.nt; g.neutral = 1; g.colid0 = s.colid0; g.sh0 = s.sh0;
            aw.segs.seg[aw.segs.nseg++] = g;
        }
        aw.cols = C; aw.K = L.K; aw.N = L.N; aw.key = pkey; aw.step = m.step; aw.sup = sup; aw.d = aw_d_.p;
        if (m.sup) aw_export_kernel<real, true><<<L.K, BLOCK, 0, stream_>>>(aw);
        else aw_export_kernel<real, false><<<L.K, BLOCK, 0, stream_>>>(aw);
        ++launches;
    }
    XchgWaitArgs xw{};
    XchgPostArgs xp{};
    if (xchg_on_ && L.world > 1) {
        // one-shot all-reduce over NVLink peer memory, completed inside shared_kernel
        ++xchg_seq_;
        xp.sums = sums_.p; xp.P = (int)sums_.n; xp.world = L.world; xp.rank = L.rank;
        xp.parity = (int)(xchg_seq_ & 1ull); xp.seq = xchg_seq_;
        for (int r = 0; r < L.world; ++r) {
            xp.peer_buf[r] = reinterpret_cast<double *>(xchg_peer_mem_[r]);
            xp.peer_flag[r] = reinterpret_cast<unsigned long long *>(static_cast<char *>(xchg_peer_mem_[r]) + xchg_flag_off_);
        }
        if (!use_tail) {
            xchg_post_kernel<<<1, 256, 0, stream_>>>(xp);
            ++launches;
        }
        xw.buf = reinterpret_cast<const double *>(xchg_mem_);
        xw.flag = reinterpret_cast<const unsigned long long *>(static_cast<char *>(xchg_mem_) + xchg_flag_off_);
        xw.P = (int)sums_.n; xw.world = L.world; xw.parity = xp.parity; xw.seq = xchg_seq_; xw.err = xchg_err_.p;
    } else if (comm_) {
        int rc = NcclApi::get().AllReduce(sums_.p, sums_.p, sums_.n, kNcclFloat64, kNcclSum, comm_, stream_);
        if (rc != 0) throw std::runtime_error(std::string("ncclAllReduce: ") + NcclApi::get().GetErrorString(rc));
    }
    // ---- shared latents
    SharedArgs<real> sa{};
    {
        sa.R = L.R; sa.K = L.K; sa.tmax = L.tmax; sa.nst = L.nst;
        for (int r = 0; r < L.R; ++r) { sa.nt[r] = L.nt[r]; sa.sh0[r] = L.sh0[r]; }
        sa.n_neutral = (double)L.N;
        sa.sums = sums_.p; sa.xchg = xw; sa.sh_th = sh_th_.p; sa.sh_acc = sh_acc_.p;
        if (sh_ring_.p) {               // n + 1 slots indexed by the step itself (bb_aux_kernels.cuh)
            const uint32_t n1 = (uint32_t)opt_.n + 1u;
            sa.sh_ring_rd = sh_ring_.p + (size_t)((m.step + 1u) % n1) * 2 * L.nst;
            sa.sh_ring_wr = sh_ring_.p + (size_t)(m.step % n1) * 2 * L.nst;
        }
        sa.ring_writer = 1;
        sa.sh_pr = sh_pr_.p;
        sa.key = pkey; sa.step = m.step;
        sa.eps_sh = m.sup ? sup_sh_.p : nullptr; sa.z_direct = m.z_direct ? 1 : 0;
        sa.ctx = ctx_.p; sa.scratch = sh_scratch_.p;
        sa.gout = m.gout ? sh_gout_.p : nullptr;
        sa.dump = m.dump ? dump_sh_.p : nullptr;
        sa.elbo_sh = m.want_elbo ? elbo_sh_.p : nullptr;
        sa.opt = opt_args<double>(m.update);
        sa.leader = L.rank == 0 ? 1 : 0;
        sa.aw_d = aw_d_.p; sa.aw_zs = aw_zs_.p; sa.aw_tmp = aw_tmp_.p; sa.aw_N = L.N;
        if (use_tail)
            tail_kernel<real><<<cdiv((long long)nwarps * 32, 256) + 1, 256, tail_smem, stream_>>>(ra, L.R, xp, sa, ticket_.p,
                                                                                             (int)sums_.n);
        else
            shared_kernel<real><<<1, 256, 0, stream_>>>(sa);
        ++launches;
    }
    // ---- pass 2
    if (tev_pos_ >= 0) BB_CUDA(cudaEventRecord(tev_[tev_pos_++], stream_));
    for (Group &g : groups_) {
        P2Args<real> a{};
        a.segs = g.p2segs; a.cols = C; a.K = L.K; a.ne = L.E;
        for (int t = 0; t < MAX_NT_DYN; ++t) a.env_of_t[t] = g.env_of_t[t];
        a.key = pkey; a.step = m.step;
        a.hy_zeps = zeps_.p; a.H = L.H; a.hz_k = hz_k_; a.hz_h = hz_h_;
        a.ctx = ctx_.p; a.tmax_ctx = L.tmax;
        a.opt = opt_args<real>(m.update);
        a.gout_lam = m.gout ? gout_lam_.p : nullptr; a.gout_bc = m.gout ? gout_bc_.p : nullptr;
        a.hcontrib = hcontrib_.p;
        // the supplied-noise kernels always carry the ELBO terms (ELBO = true instantiation only)
        const bool elbo = m.want_elbo || m.sup;
        a.epart = elbo ? epart_.p + (size_t)g.epart_off * (L.K + 1) : nullptr;
        a.sup = sup;
        a.stage_pr = (L.lam_pr_matrix || L.bc_pr_matrix) ? 1 : 0;
        // smem is sized for the ring by size_pass2(); stage it only when the ring exists and is updated
        a.stage_ring = (opt_.kind == BB_OPT_TRUNCATED_ADAGRAD && lam_ring_.p && m.update && g.p2stage_ring) ? 1 : 0;
        a.l2_ring = (opt_.kind == BB_OPT_TRUNCATED_ADAGRAD && lam_ring_.p && m.update && !(a.stage_ring && g.p2stage_acc)) ? 1 : 0;
        a.stage_acc = g.p2stage_acc; a.nbuf = g.p2nbuf;
        a.abort = (xchg_on_ && L.world > 1) ? xchg_err_.p : nullptr;
        a.aw_zs = aw_zs_.p; a.aw_N = L.N;
        if (m.stepk) {
            // the packed step kernel: one step behind the tail (programmatic dependent launch), or -- persistent,
            // cooperative -- m.nsteps steps with the tails of steps 2.. inside the kernel
            StepArgs<real> sk{};
            sk.segs = g.stsegs; sk.cols = col_arrays(true); sk.K = L.K; sk.P = (int)sums_.n;
            for (int t = 0; t < MAX_NT_DYN; ++t) sk.env_of_t[t] = g.env_of_t[t];
            sk.key = pkey; sk.step = m.step; sk.nsteps = m.nsteps;
            sk.ctx = ctx_.p;
            sk.opt = opt_args<real>(true);
            sk.stage_pr = (L.lam_pr_matrix || L.bc_pr_matrix) ? 1 : 0;
            const bool trunc = opt_.kind == BB_OPT_TRUNCATED_ADAGRAD && lam_ring_.p;
            sk.stage_ring = (trunc && g.step_stage_ring) ? 1 : 0;
            sk.l2_ring = (trunc && !g.step_stage_ring) ? 1 : 0;
            sk.nbuf = g.step_nbuf; sk.stage_acc = g.step_stage_acc;
            sk.acc_rows = g.step_acc_rows; sk.tail_scratch = (int)sh_scratch_.n;
            sk.abort = (xchg_on_ && L.world > 1) ? xchg_err_.p : nullptr;
            sk.xpart = xpart_.p;
            sk.ring_n = trunc ? opt_.n : 0; sk.ring_slot = ring_slot_;
            if (m.nsteps > 1) {
                sk.sync = step_sync_.p; sk.gpart = gpart_.p; sk.gsize = step_gsize_;
                sk.ngroups = cdiv(g.stblocks, step_gsize_);
                sk.xp.P = (int)sums_.n; sk.xp.world = L.world; sk.xp.rank = L.rank;
                sk.xp.seq = xchg_seq_ + 1; sk.xp.parity = 0; sk.xp.sums = nullptr;
                for (int r = 0; r < L.world; ++r) {
                    sk.xp.peer_buf[r] = reinterpret_cast<double *>(xchg_peer_mem_[r]);
                    sk.xp.peer_flag[r] = reinterpret_cast<unsigned long long *>(static_cast<char *>(xchg_peer_mem_[r]) + xchg_flag_off_);
                }
                sk.xbuf = reinterpret_cast<const double *>(xchg_mem_);
                sk.xflag = reinterpret_cast<const unsigned long long *>(static_cast<char *>(xchg_mem_) + xchg_flag_off_);
                xchg_seq_ += (unsigned long long)(m.nsteps - 1);
                sk.sa = sa;
                sk.sa.xchg = XchgWaitArgs{};
                sk.sa.sh_ring_rd = sh_ring_.p; sk.sa.sh_ring_wr = sh_ring_.p;      // ring base: the kernel picks the slots
                sk.sa.eps_sh = eps_steps_.p; sk.sa.z_direct = 0;
                sk.sa.gout = nullptr; sk.sa.dump = nullptr; sk.sa.elbo_sh = nullptr;
                const int ne = m.nsteps * L.K * 2 * L.nst;
                shared_noise_steps_kernel<<<cdiv(ne, 128), 128, 0, stream_>>>(pkey, m.step, m.nsteps, L.K, 2 * L.nst, eps_steps_.p);
                ++launches;
                void *kargs[] = {(void *)&sk};
                BB_CUDA(cudaLaunchCooperativeKernel((const void *)g.stepk, dim3(g.stblocks), dim3(BLOCK), kargs, g.step_smem, stream_));
            } else {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(g.stblocks); cfg.blockDim = dim3(BLOCK); cfg.dynamicSmemBytes = g.step_smem; cfg.stream = stream_;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = getenv("BB_NO_PDL") ? 0 : 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                BB_CUDA(cudaLaunchKernelEx(&cfg, g.stepk, sk));
            }
        } else if (m.fuse) {
            // the reducer has consumed this step's partials (stream order): the fused kernel overwrites them
            a.part = part_.p + (size_t)g.epart_off * L.K * (3 * L.tmax);
            a.pv = g.pv; a.acc_slots = g.facc_slots;
            // programmatic dependent launch: the CTAs become resident while the tail kernel is still running and
            // do everything that does not need its output (direction table, accumulator zeroing, first tile)
            // before griddepcontrol.wait
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(g.p2blocks); cfg.blockDim = dim3(BLOCK); cfg.dynamicSmemBytes = g.fsmem; cfg.stream = stream_;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = getenv("BB_NO_PDL") ? 0 : 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            BB_CUDA(cudaLaunchKernelEx(&cfg, g.ks.pass2_fused, a));
        } else {
            // also a programmatic dependent of the kernel ahead of it (the tail kernel triggers early): its prologue
            // -- direction table, first tile of theta -- runs beside the shared-latent phases (hierarchical models:
            // -2 us per step)
            const KernelSet<real> &ks = m.sup ? g.ks_sup : g.ks;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(g.p2blocks); cfg.blockDim = dim3(BLOCK); cfg.stream = stream_;
            cfg.dynamicSmemBytes = elbo ? g.p2smem_elbo : g.p2smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = getenv("BB_NO_PDL") ? 0 : 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            BB_CUDA(cudaLaunchKernelEx(&cfg, elbo ? ks.pass2_elbo : ks.pass2, a));
        }
        ++launches;
    }
    if (tev_pos_ >= 0) BB_CUDA(cudaEventRecord(tev_[tev_pos_++], stream_));
    if (L.hier && L.H > 0) {
        // many members per hyper latent (genotypes): warp-per-latent gather first
        if (!m.dump && csr_mem_.n >= (size_t)8 * L.H) {
            hsum_.ensure((size_t)L.H);
            ha.hsum = hsum_.p;
            hyper_gather_kernel<real><<<cdiv((long long)L.H * 32, BLOCK), BLOCK, 0, stream_>>>(ha);
            ++launches;
        }
        hyper_update_kernel<real><<<hyblocks_, BLOCK, (size_t)(L.K + 1) * BLOCK * sizeof(double), stream_>>>(ha);
        ++launches;
    }
    if (m.want_elbo) {
        elbo_sum_kernel<<<L.K + 1, 128, 0, stream_>>>(epart_.p, p2blocks_total_, (L.hier && L.H > 0) ? hy_epart_.p : nullptr,
                                                     (L.hier && L.H > 0) ? hyblocks_ : 0, elbo_sh_.p, L.K, elbo_out_.p);
        ++launches;
    }
    BB_CUDA(cudaGetLastError());
}

// ELBO from elbo_out_ (device): (1/K) sum_k (logp_k) + sum log sigma + D (1 + log 2 pi) / 2
template <typename real> void Engine<real>::finish_elbo_partial(double *out) {
    BB_CUDA(cudaMemcpyAsync(out, elbo_out_.p, sizeof(double) * (L.K + 1), cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
}

template <typename real> double Engine<real>::finish_elbo(double *logp_k) {
    std::vector<double> out(L.K + 1);
    if (comm_) {
        int rc = NcclApi::get().AllReduce(elbo_out_.p, elbo_out_.p, L.K + 1, kNcclFloat64, kNcclSum, comm_, stream_);
        if (rc != 0) throw std::runtime_error("ncclAllReduce (elbo) failed");
    } else if (xchg_on_ && L.world > 1 && !raw_elbo_) {
        peer_allreduce(elbo_out_.p, (size_t)L.K + 1);
    }
    BB_CUDA(cudaMemcpyAsync(out.data(), elbo_out_.p, sizeof(double) * (L.K + 1), cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    // raw_elbo_ (one shard of a single-process multi-GPU handle): no constants, the owner adds them once after summing
    const double cst = raw_elbo_ ? 0.0 : L.logp_const;
    double mean = 0.0;
    for (int k = 0; k < L.K; ++k) {
        const double lp = out[k] + cst;
        if (logp_k) logp_k[k] = lp;
        mean += lp / L.K;
    }
    return mean + out[L.K] + (raw_elbo_ ? 0.0 : 0.5 * (double)L.D * (1.0 + std::log(2.0 * M_PI)));
}

template <typename real>
void Engine<real>::logjoint_grad(const double *x, int K, int eps_is_noise, double *logp, double *grad) {
    if (K != L.K) throw std::runtime_error("n_samples must equal the handle's samples_per_step");
    ensure_supplied(true);
    upload_supplied(x, K);
    RunMode m; m.sup = true; m.z_direct = !eps_is_noise; m.want_elbo = true; m.dump = true;
    run_pipeline(m);
    finish_elbo(logp);
    // gather the per-sample gradients to reference order
    hostvec_b_.ensure((size_t)K * L.D);
    BB_CUDA(cudaMemsetAsync(hostvec_b_.p, 0, sizeof(double) * K * L.D, stream_));
    auto run = [&](const real *src, const int *map, size_t n) {
        if (!n) return;
        gather_rows_kernel<real><<<cdiv(n, 256), 256, 0, stream_>>>(src, map, (long long)n, K, hostvec_b_.p, L.D);
        ++launches;
    };
    run(dump_lam_.p, map_lam_.p, lam_th_.n);
    run(dump_bc_.p, map_bc_.p, bc_th_.n);
    if (L.hier) run(dump_hy_.p, map_hy_.p, hy_th_.n);
    if (L.rank == 0)
        BB_CUDA(cudaMemcpy2DAsync(hostvec_b_.p, sizeof(double) * L.D, dump_sh_.p, sizeof(double) * 2 * L.nst,
                                  sizeof(double) * 2 * L.nst, K, cudaMemcpyDeviceToDevice, stream_));
    BB_CUDA(cudaMemcpyAsync(grad, hostvec_b_.p, sizeof(double) * K * L.D, cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    BB_CUDA(cudaGetLastError());
}

template <typename real> void Engine<real>::elbo_grad(const double *eps, long long step, double *elbo, double *grad) {
    gout_lam_.ensure(lam_th_.n); gout_bc_.ensure(std::max<size_t>(bc_th_.n, 1));
    if (L.hier) gout_hy_.ensure(std::max<size_t>(hy_th_.n, 1));
    RunMode m; m.want_elbo = true; m.gout = true; m.step = (uint32_t)step;
    if (eps) { ensure_supplied(false); upload_supplied(eps, L.K); m.sup = true; }
    run_pipeline(m);
    const double e = finish_elbo(nullptr);
    if (elbo) *elbo = e;
    if (grad) {
        hostvec_a_.ensure(L.D); hostvec_b_.ensure(L.D);
        gather_to_ref(gout_lam_.p, gout_bc_.p, gout_hy_.p, sh_gout_.p, hostvec_a_.p, hostvec_b_.p, false);
        BB_CUDA(cudaMemcpyAsync(grad, hostvec_a_.p, sizeof(double) * L.D, cudaMemcpyDeviceToHost, stream_));
        BB_CUDA(cudaMemcpyAsync(grad + L.D, hostvec_b_.p, sizeof(double) * L.D, cudaMemcpyDeviceToHost, stream_));
        BB_CUDA(cudaStreamSynchronize(stream_));
    }
    BB_CUDA(cudaGetLastError());
}

template <typename real> void Engine<real>::get_noise(long long step, double *eps) {
    const PhiloxKey pkey = philox_key(seed_);
    const size_t n = (size_t)L.K * L.D;
    hostvec_a_.ensure(n);
    BB_CUDA(cudaMemsetAsync(hostvec_a_.p, 0, sizeof(double) * n, stream_));
    SegList sl{};
    for (const HostSeg &s : L.segs) {
        Seg g{}; g.col0 = s.col0; g.ncol = s.ncol; g.rep = s.rep; g.nt = s.nt; g.neutral = s.neutral; g.colid0 = s.colid0;
        sl.seg[sl.nseg++] = g;
    }
    for (int t = 0; t < L.tmax; ++t) {
        noise_columns_kernel<real><<<cdiv(L.cpad, 128), 128, 0, stream_>>>(sl, L.cpad, t, 0, col_id_.p,
            map_lam_.p + (size_t)t * L.cpad, L.K, L.D, (uint32_t)step, pkey, hostvec_a_.p);
        ++launches;
    }
    for (int j = 0; j < L.nj; ++j) {
        noise_columns_kernel<real><<<cdiv(L.cpad, 128), 128, 0, stream_>>>(sl, L.cpad, j, 1, col_id_.p,
            map_bc_.p + (size_t)j * L.cpad, L.K, L.D, (uint32_t)step, pkey, hostvec_a_.p);
        ++launches;
    }
    if (L.hier && L.H > 0) {
        noise_stream_kernel<real><<<cdiv(L.H, 128), 128, 0, stream_>>>(STREAM_HYPER, L.hy_gid0, map_hy_.p, L.H, L.K,
                                                                       L.D, (uint32_t)step, pkey, 0, hostvec_a_.p);
        ++launches;
    }
    if (L.nst) {
        noise_stream_kernel<real><<<cdiv(2 * L.nst, 128), 128, 0, stream_>>>(STREAM_SHARED, 0u, map_sh_.p, 2 * L.nst,
                                                                             L.K, L.D, (uint32_t)step, pkey, 1,
                                                                             hostvec_a_.p);
        ++launches;
    }
    BB_CUDA(cudaMemcpyAsync(eps, hostvec_a_.p, sizeof(double) * n, cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    BB_CUDA(cudaGetLastError());
}

// fp32 TruncatedADAGrad: rebuild the window sums of the column and hyper latents from the ring at window wraps -- the
// first four wraps (the large first gradients leave the window there) and every eighth after them: n ring reads per
// latent, amortised below 1 % of a step.  Depends on the step counter only, so split, resumed and sharded runs re-sum
// at the same steps.  BB_NO_RESUM=1: off (the behaviour until the last build of round 2).
template <typename real> bool Engine<real>::resum_active() const {
    return std::is_same<real, float>::value && opt_.kind == BB_OPT_TRUNCATED_ADAGRAD && lam_ring_.p != nullptr &&
           !getenv("BB_NO_RESUM");
}
template <typename real> void Engine<real>::maybe_resum() {
    if (!resum_active() || step_count <= 0 || step_count % opt_.n != 0) return;
    const long long wraps = step_count / opt_.n;
    if (wraps > 4 && wraps % 8 != 0) return;
    auto run = [&](const DBuf<r2> &ring, DBuf<r2> &acc) {
        if (!ring.p || !acc.p || acc.n == 0) return;
        ring_resum_kernel<real><<<cdiv((long long)acc.n, 256), 256, 0, stream_>>>(ring.p, acc.p, (long long)acc.n, opt_.n);
        ++launches;
    };
    run(lam_ring_, lam_acc_); run(bc_ring_, bc_acc_);
    if (L.hier) run(hy_ring_, hy_acc_);
    BB_CUDA(cudaGetLastError());
}

template <typename real> void Engine<real>::step(int n, double *trace) {
    if (!opt_ready_) set_optimizer(opt_);
    if (trace) trace_.ensure((size_t)n * (L.K + 1));
    const bool use_stepk = stepk_ok_ && trace == nullptr;
    // persistent launches need the merged tail kernel ahead of them and, multi-GPU, the peer-memory exchange
    const bool can_persist = use_stepk && persist_chunk_ > 1 && !(comm_ && !xchg_on_) && !getenv("BB_NO_TAIL");
    bool persisted = false;
    for (int i = 0; i < n;) {
        maybe_resum();
        RunMode m; m.update = true; m.want_elbo = trace != nullptr; m.step = (uint32_t)step_count;
        if (use_stepk) {
            m.fuse = true; m.stepk = true;
            m.have_xpart = part_step_ == step_count && part_is_x_;
            m.have_part = part_step_ == step_count && !part_is_x_;
            m.nsteps = can_persist ? std::min(persist_chunk_, n - i) : 1;
            // a persistent launch ends at the window wrap, where the sums are rebuilt between launches
            if (resum_active()) m.nsteps = (int)std::min<long long>(m.nsteps, opt_.n - step_count % opt_.n);
            persisted = persisted || m.nsteps > 1;
        } else {
            // software-pipelined step: pass 2 of this step also produces the partial sums of the next one
            m.fuse = fused_ok_ && trace == nullptr;
            m.have_part = fused_ok_ && part_step_ == step_count && !part_is_x_;
            m.have_xpart = part_step_ == step_count && part_is_x_;
        }
        run_pipeline(m);
        part_step_ = m.fuse ? step_count + m.nsteps : -1;
        part_is_x_ = m.stepk;
        if (trace)
            BB_CUDA(cudaMemcpyAsync(trace_.p + (size_t)i * (L.K + 1), elbo_out_.p, sizeof(double) * (L.K + 1),
                                    cudaMemcpyDeviceToDevice, stream_));
        step_count += m.nsteps;
        i += m.nsteps;
        if (opt_.kind == BB_OPT_TRUNCATED_ADAGRAD) ring_slot_ = (int)(step_count % opt_.n);
    }
    // the exchange inside a kernel cannot raise a host error by itself: surface it with the call that ran the steps
    if (persisted || (xchg_on_ && L.world > 1)) check_step_sync();
    if (trace) {
        if (comm_) {
            int rc = NcclApi::get().AllReduce(trace_.p, trace_.p, (size_t)n * (L.K + 1), kNcclFloat64, kNcclSum, comm_, stream_);
            if (rc != 0) throw std::runtime_error("ncclAllReduce (trace) failed");
        } else if (xchg_on_ && L.world > 1 && !raw_elbo_) {
            peer_allreduce(trace_.p, (size_t)n * (L.K + 1));
        }
        std::vector<double> h((size_t)n * (L.K + 1));
        BB_CUDA(cudaMemcpyAsync(h.data(), trace_.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, stream_));
        BB_CUDA(cudaStreamSynchronize(stream_));
        for (int i = 0; i < n; ++i) {
            double mean = 0.0;
            for (int k = 0; k < L.K; ++k) mean += (h[(size_t)i * (L.K + 1) + k] + (raw_elbo_ ? 0.0 : L.logp_const)) / L.K;
            trace[i] = mean + h[(size_t)i * (L.K + 1) + L.K] + (raw_elbo_ ? 0.0 : 0.5 * (double)L.D * (1.0 + std::log(2.0 * M_PI)));
        }
    }
}

template <typename real> void Engine<real>::step_with_noise(const double *eps) {
    if (!opt_ready_) set_optimizer(opt_);
    ensure_supplied(false);
    upload_supplied(eps, L.K);
    maybe_resum();
    RunMode m; m.sup = true; m.update = true; m.step = (uint32_t)step_count;
    run_pipeline(m);
    ++step_count;
    if (opt_.kind == BB_OPT_TRUNCATED_ADAGRAD) ring_slot_ = (int)(step_count % opt_.n);
    BB_CUDA(cudaStreamSynchronize(stream_));
}

template <typename real> void Engine<real>::time_steps(int n, float *ms_total, float *ms_pass1, float *ms_pass2) {
    // CUDA-event timing on the launching stream: whole region, plus per-step brackets around the
    // pass-1 and pass-2 column kernels (with ragged T the bracket spans the group launches).
    if (!opt_ready_) set_optimizer(opt_);
    while ((int)tev_.size() < 4 * n + 4) {
        cudaEvent_t e;
        BB_CUDA(cudaEventCreate(&e));
        tev_.push_back(e);
    }
    tev_pos_ = 0;
    BB_CUDA(cudaEventRecord(ev_[0], stream_));
    try { step(n, nullptr); } catch (...) { tev_pos_ = -1; throw; }
    BB_CUDA(cudaEventRecord(ev_[1], stream_));
    BB_CUDA(cudaEventSynchronize(ev_[1]));
    BB_CUDA(cudaEventElapsedTime(ms_total, ev_[0], ev_[1]));
    float p1 = 0.f, p2 = 0.f;
    const int nbr = tev_pos_ / 4;         // one bracket set per launch sequence (a persistent launch covers many steps)
    tev_pos_ = -1;
    for (int i = 0; i < nbr; ++i) {
        float a = 0.f, b = 0.f;
        BB_CUDA(cudaEventElapsedTime(&a, tev_[4 * i + 0], tev_[4 * i + 1]));
        BB_CUDA(cudaEventElapsedTime(&b, tev_[4 * i + 2], tev_[4 * i + 3]));
        p1 += a; p2 += b;
    }
    if (ms_pass1) *ms_pass1 = p1;
    if (ms_pass2) *ms_pass2 = p2;
}

// ------------------------------------------------------------------ state
// [step_count, ring_slot, mu[D], omega[D], acc_mu[D], acc_omega[D]] and, for TruncatedADAGrad, the window itself:
// n slots of (g_mu^2[D], g_omega^2[D]) in reference order, then the shared latents' ring ((n + 1) x 2 nst pairs, raw).
// A restored run continues exactly like an uninterrupted one (tests/test_gpu_models.py).
template <typename real> long long Engine<real>::state_size() const {
    long long n = 2 + 4 * L.D;
    if (opt_.kind == BB_OPT_TRUNCATED_ADAGRAD) n += 2LL * opt_.n * L.D + 4LL * L.nst * (opt_.n + 1);
    return n;
}

template <typename real> void Engine<real>::get_state(double *s) {
    if (!opt_ready_) set_optimizer(opt_);
    s[0] = (double)step_count; s[1] = (double)ring_slot_;
    get_params(s + 2, s + 2 + L.D, false);
    hostvec_a_.ensure(L.D); hostvec_b_.ensure(L.D);
    gather_to_ref(lam_acc_.p, bc_acc_.p, hy_acc_.p, sh_acc_.p, hostvec_a_.p, hostvec_b_.p, false);
    BB_CUDA(cudaMemcpyAsync(s + 2 + 2 * L.D, hostvec_a_.p, sizeof(double) * L.D, cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaMemcpyAsync(s + 2 + 3 * L.D, hostvec_b_.p, sizeof(double) * L.D, cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    if (opt_.kind != BB_OPT_TRUNCATED_ADAGRAD) return;
    double *r = s + 2 + 4 * L.D;
    for (int slot = 0; slot < opt_.n; ++slot) {
        gather_to_ref(lam_ring_.p + (size_t)slot * lam_th_.n, bc_ring_.p + (size_t)slot * bc_th_.n,
                      hy_ring_.p ? hy_ring_.p + (size_t)slot * hy_th_.n : nullptr, nullptr, hostvec_a_.p, hostvec_b_.p, false);
        BB_CUDA(cudaMemcpyAsync(r + (size_t)slot * 2 * L.D, hostvec_a_.p, sizeof(double) * L.D, cudaMemcpyDeviceToHost, stream_));
        BB_CUDA(cudaMemcpyAsync(r + (size_t)slot * 2 * L.D + L.D, hostvec_b_.p, sizeof(double) * L.D, cudaMemcpyDeviceToHost, stream_));
        BB_CUDA(cudaStreamSynchronize(stream_));
    }
    double *rs = r + (size_t)2 * opt_.n * L.D;
    const size_t nsh = (size_t)4 * L.nst * (opt_.n + 1);
    if (L.rank == 0) BB_CUDA(cudaMemcpy(rs, sh_ring_.p, sizeof(double) * nsh, cudaMemcpyDeviceToHost));
    else std::fill(rs, rs + nsh, 0.0);           // replicated on every shard: reported once
}

template <typename real> void Engine<real>::set_state(const double *s) {
    if (!opt_ready_) set_optimizer(opt_);
    set_params(s + 2, s + 2 + L.D);
    hostvec_a_.ensure(L.D); hostvec_b_.ensure(L.D);
    auto scatter = [&](const double *x, const double *y, r2 *lam, r2 *bc, r2 *hy) {
        BB_CUDA(cudaMemcpyAsync(hostvec_a_.p, x, sizeof(double) * L.D, cudaMemcpyHostToDevice, stream_));
        BB_CUDA(cudaMemcpyAsync(hostvec_b_.p, y, sizeof(double) * L.D, cudaMemcpyHostToDevice, stream_));
        auto run = [&](r2 *dst, const int *map, size_t n) {
            if (!n || !dst) return;
            scatter_pairs_kernel<real><<<cdiv(n, 256), 256, 0, stream_>>>(dst, map, (long long)n, hostvec_a_.p,
                                                                          hostvec_b_.p, 0.0, 0.0);
            ++launches;
        };
        run(lam, map_lam_.p, lam_th_.n);
        run(bc, map_bc_.p, bc_th_.n);
        if (L.hier) run(hy, map_hy_.p, hy_th_.n);
        BB_CUDA(cudaStreamSynchronize(stream_));
    };
    const double *am = s + 2 + 2 * L.D, *ao = s + 2 + 3 * L.D;
    scatter(am, ao, lam_acc_.p, bc_acc_.p, hy_acc_.p);
    std::vector<double2> sh(2 * L.nst);
    for (int i = 0; i < 2 * L.nst; ++i) sh[i] = make_double2(am[i], ao[i]);
    BB_CUDA(cudaMemcpyAsync(sh_acc_.p, sh.data(), sizeof(double2) * sh.size(), cudaMemcpyHostToDevice, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    if (opt_.kind == BB_OPT_TRUNCATED_ADAGRAD) {
        const double *r = s + 2 + 4 * L.D;
        for (int slot = 0; slot < opt_.n; ++slot)
            scatter(r + (size_t)slot * 2 * L.D, r + (size_t)slot * 2 * L.D + L.D, lam_ring_.p + (size_t)slot * lam_th_.n,
                    bc_ring_.p + (size_t)slot * bc_th_.n, hy_ring_.p ? hy_ring_.p + (size_t)slot * hy_th_.n : nullptr);
        BB_CUDA(cudaMemcpy(sh_ring_.p, r + (size_t)2 * opt_.n * L.D, sizeof(double) * 4 * L.nst * (opt_.n + 1),
                           cudaMemcpyHostToDevice));
    }
    step_count = (long long)s[0];
    ring_slot_ = (int)s[1];
    part_step_ = -1;
    hy_zeps_step_ = -1;
}

// ------------------------------------------------------------------ comm
template <typename real> void Engine<real>::comm_init(const char id[128]) {
    if (L.world == 1) return;
    NcclApi &api = NcclApi::get();
    NcclApi::UniqueId uid;
    std::memcpy(uid.internal, id, 128);
    BB_CUDA(cudaSetDevice(device_));
    int rc = api.CommInitRank(&comm_, L.world, uid, L.rank);
    if (rc != 0) throw std::runtime_error(std::string("ncclCommInitRank: ") + api.GetErrorString(rc));
    if (!getenv("BB_NO_P2P")) setup_peer_exchange();
    size_pass2();          // the persistent step kernel needs the peer-memory exchange
}

// Exchange buffer [2][world][P] doubles + flags [2][world], shared with every peer through CUDA IPC
// (handles all-gathered over the NCCL communicator just created).  Any failure leaves the NCCL
// all-reduce in place.
template <typename real> void Engine<real>::setup_peer_exchange() {
    if (L.world > MAX_WORLD) return;
    NcclApi &api = NcclApi::get();
    alloc_xchg();
    cudaIpcMemHandle_t mine;
    const bool dbg = getenv("BB_DEBUG") != nullptr;
    cudaError_t ge = cudaIpcGetMemHandle(&mine, xchg_mem_);
    int ok = ge == cudaSuccess ? 1 : 0;
    if (!ok) { cudaGetLastError(); if (dbg) fprintf(stderr, "[bb rank %d] cudaIpcGetMemHandle: %s\n", L.rank, cudaGetErrorString(ge)); }
    DBuf<unsigned char> send, recv;
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    send.alloc(rec); recv.alloc(rec * L.world);
    std::vector<unsigned char> h(rec, 0);
    std::memcpy(h.data(), &mine, sizeof(mine));
    h[sizeof(mine)] = (unsigned char)ok;
    BB_CUDA(cudaMemcpyAsync(send.p, h.data(), rec, cudaMemcpyHostToDevice, stream_));
    int rc = api.AllGather(send.p, recv.p, rec, /*ncclUint8*/ 1, comm_, stream_);
    if (rc != 0) throw std::runtime_error(std::string("ncclAllGather: ") + api.GetErrorString(rc));
    std::vector<unsigned char> all(rec * L.world);
    BB_CUDA(cudaMemcpyAsync(all.data(), recv.p, all.size(), cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    for (int r = 0; r < L.world; ++r) ok &= all[r * rec + sizeof(mine)];
    xchg_peer_mem_.assign(L.world, nullptr);
    for (int r = 0; r < L.world && ok; ++r) {
        if (r == L.rank) { xchg_peer_mem_[r] = xchg_mem_; continue; }
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, all.data() + r * rec, sizeof(hd));
        cudaError_t oe = cudaIpcOpenMemHandle(&xchg_peer_mem_[r], hd, cudaIpcMemLazyEnablePeerAccess);
        if (oe != cudaSuccess) {
            if (dbg) fprintf(stderr, "[bb rank %d] cudaIpcOpenMemHandle(peer %d): %s\n", L.rank, r, cudaGetErrorString(oe));
            cudaGetLastError();
            xchg_peer_mem_[r] = nullptr;
            ok = 0;
        }
    }
    // every rank must take the same path: agree on success with one more (tiny) collective
    DBuf<int> flag;
    flag.alloc(1);
    BB_CUDA(cudaMemcpyAsync(flag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, stream_));
    rc = api.AllReduce(flag.p, flag.p, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, comm_, stream_);
    if (rc != 0) throw std::runtime_error(std::string("ncclAllReduce: ") + api.GetErrorString(rc));
    BB_CUDA(cudaMemcpyAsync(&ok, flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    xchg_on_ = ok != 0;
    if (dbg) fprintf(stderr, "[bb rank %d] peer-memory exchange %s\n", L.rank, xchg_on_ ? "enabled" : "disabled (NCCL all-reduce)");
}

// exchange buffer [2][world][P] doubles + flags [2][world] (bb_aux_kernels.cuh "peer-memory exchange")
template <typename real> void Engine<real>::alloc_xchg() {
    if (xchg_mem_) return;
    const size_t P = sums_.n;
    xchg_flag_off_ = ((size_t)2 * L.world * xchg_slot_doubles((int)P) * sizeof(double) + 255) / 256 * 256;   // spread slots, bb_aux_kernels.cuh
    // behind the step exchange: an auxiliary region [2][world][kAuxP] + flags for short all-reduces (ELBO terms)
    xchg_aux_off_ = (xchg_flag_off_ + (size_t)2 * L.world * sizeof(unsigned long long) + 255) / 256 * 256;
    xchg_aux_flag_off_ = xchg_aux_off_ + (size_t)2 * L.world * kAuxP * sizeof(double);
    const size_t bytes = xchg_aux_flag_off_ + (size_t)2 * L.world * sizeof(unsigned long long);
    BB_CUDA(cudaMalloc(&xchg_mem_, bytes));
    BB_CUDA(cudaMemset(xchg_mem_, 0, bytes));
    if (!xchg_err_.p) xchg_err_.alloc(1);
}

template <typename real> void Engine<real>::peer_allreduce(double *dev, size_t n) {
    char *base = static_cast<char *>(xchg_mem_);
    for (size_t off = 0; off < n; off += kAuxP) {
        const int m = (int)std::min<size_t>(kAuxP, n - off);
        ++aux_seq_;
        XchgPostArgs xp{};
        xp.P = kAuxP; xp.world = L.world; xp.rank = L.rank; xp.parity = (int)(aux_seq_ & 1ull); xp.seq = aux_seq_;
        for (int r = 0; r < L.world; ++r) {
            char *pb = static_cast<char *>(xchg_peer_mem_[r]);
            xp.peer_buf[r] = reinterpret_cast<double *>(pb + xchg_aux_off_);
            xp.peer_flag[r] = reinterpret_cast<unsigned long long *>(pb + xchg_aux_flag_off_);
        }
        XchgWaitArgs xw{};
        xw.buf = reinterpret_cast<const double *>(base + xchg_aux_off_);
        xw.flag = reinterpret_cast<const unsigned long long *>(base + xchg_aux_flag_off_);
        xw.P = kAuxP; xw.world = L.world; xw.parity = xp.parity; xw.seq = aux_seq_; xw.err = xchg_err_.p;
        peer_allreduce_kernel<<<1, 256, 0, stream_>>>(dev + off, m, xp, xw);
        ++launches;
    }
}

template <typename real> void Engine<real>::peer_handle(char out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    alloc_xchg();
    cudaIpcMemHandle_t h;
    BB_CUDA(cudaIpcGetMemHandle(&h, xchg_mem_));
    std::memcpy(out, &h, 64);
}

template <typename real> void Engine<real>::peer_attach(const char *handles, int n) {
    if (n != L.world) throw std::runtime_error("bb_peer_attach: need one handle per rank (world = " + std::to_string(L.world) + ")");
    if (L.world > MAX_WORLD) throw std::runtime_error("at most 16 ranks");
    if (xchg_on_) throw std::runtime_error("bb_peer_attach: the exchange of this handle is already wired");
    alloc_xchg();
    xchg_peer_mem_.assign(L.world, nullptr);
    for (int r = 0; r < L.world; ++r) {
        if (r == L.rank) { xchg_peer_mem_[r] = xchg_mem_; continue; }
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, handles + (size_t)r * 64, 64);
        BB_CUDA(cudaIpcOpenMemHandle(&xchg_peer_mem_[r], hd, cudaIpcMemLazyEnablePeerAccess));
    }
    xchg_on_ = true;
    size_pass2();
}

// single-process multi-GPU: every device's exchange buffer, mapped by peer access (no IPC, no NCCL)
template <typename real> void Engine<real>::attach_peers(const std::vector<void *> &bufs) {
    xchg_peer_mem_ = bufs;
    xchg_on_ = true;
    peers_local_ = true;
    size_pass2();
}

// after the steps of one call: did an exchange (tail kernel or persistent step kernel) give up waiting for a peer?
template <typename real> void Engine<real>::check_step_sync() {
    int err = 0, serr = 0;
    if (xchg_err_.p) BB_CUDA(cudaMemcpyAsync(&err, xchg_err_.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    if (step_sync_.p)
        BB_CUDA(cudaMemcpyAsync(&serr, reinterpret_cast<char *>(step_sync_.p) + offsetof(StepSync, err), sizeof(int),
                                cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    BB_CUDA(cudaGetLastError());
    if (err || serr) {
        // the tickets of an aborted persistent launch are mid-cycle: reset the control block, drop the pipelined sums
        if (step_sync_.p) BB_CUDA(cudaMemset(step_sync_.p, 0, sizeof(StepSync)));
        if (xchg_err_.p) BB_CUDA(cudaMemset(xchg_err_.p, 0, sizeof(int)));
        part_step_ = -1;
    hy_zeps_step_ = -1;
        throw std::runtime_error("peer-memory exchange timed out: a rank did not post its partial sums; the steps of "
                                 "this call are incomplete");
    }
}

template <typename real> void Engine<real>::derived_fitness(int n, uint64_t seed, double *median, double *sd) {
    if (!L.hier) throw std::runtime_error("derived bc_fitness rows exist for the hierarchical models only");
    if (n < 1 || n > 12000) throw std::runtime_error("n_samples must be in [1, 12000]");
    std::vector<int> cols;
    for (const HostSeg &s : L.segs)
        if (!s.neutral)
            for (int i = 0; i < s.ncol; ++i) cols.push_back(s.col0 + i);
    const size_t nout = (size_t)L.bc_block;
    DBuf<int> dcols;
    DBuf<double> dmed, dsd;
    dcols.upload(cols);
    dmed.alloc(nout); dsd.alloc(nout);
    if (!cols.empty()) {
        DerivedArgs<real> a{};
        a.cols = dcols.p; a.ncols = (int)cols.size(); a.E = L.E; a.cpad = L.cpad; a.n = n;
        a.bc_th = bc_th_.p; a.hy_th = hy_th_.p; a.hgroup = hgroup_.p; a.map_tau = map_bc_.p; a.off_tau = L.off_bc[1];
        a.key = philox_key(seed);
        a.med = dmed.p; a.sd = dsd.p;
        const long long items = (long long)cols.size() * L.E;
        const int grid = (int)std::min<long long>(items, (long long)nsm_ * 5);
        derived_fitness_kernel<real><<<grid, DERIVED_THREADS, (size_t)n * sizeof(float), stream_>>>(a);
        ++launches;
    }
    BB_CUDA(cudaMemcpyAsync(median, dmed.p, sizeof(double) * nout, cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaMemcpyAsync(sd, dsd.p, sizeof(double) * nout, cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    BB_CUDA(cudaGetLastError());
}

template <typename real> void Engine<real>::persist_stats(double out[16]) {
    for (int i = 0; i < 16; ++i) out[i] = 0.0;
    if (!step_sync_.p) return;
    StepSync h;
    BB_CUDA(cudaMemcpy(&h, step_sync_.p, sizeof(StepSync), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) out[i] = (double)h.stat[i];
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device_);
    out[4] = (double)khz;
    out[5] = (double)h.stat[4]; out[6] = (double)h.stat[5];      // inside "sums -> context": completing the sums, shared-latent phases
    out[7] = (double)h.stat[6];                                  // the last-arriving CTA: group sum -> rank sum -> posted to the peers
    out[8] = (double)h.stat[7];                                  // ... and its column phase (the longest of the grid)
    for (int i = 0; i < 4; ++i) out[9 + i] = (double)h.stat[8 + i];  // its chain: group sum | tickets + fences | rank sum + peer stores | fence + flags
    BB_CUDA(cudaMemset(reinterpret_cast<char *>(step_sync_.p) + offsetof(StepSync, stat), 0, sizeof(h.stat)));
}

template <typename real> void Engine<real>::check_peer_exchange() {
    if (!xchg_on_) return;
    int err = 0;
    BB_CUDA(cudaMemcpyAsync(&err, xchg_err_.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    BB_CUDA(cudaStreamSynchronize(stream_));
    if (err) throw std::runtime_error("peer-memory exchange timed out: a rank did not post its partial sums");
}

}  // namespace bb
