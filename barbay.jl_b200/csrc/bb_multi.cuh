// Single-call multi-GPU: ONE handle, n_devices GPUs of this process, one blocking call per entry point --
// what a single BarBay.vi.advi() call (src/vi.jl:86-101: synchronous, one process) needs to reach more than
// one GPU (SURVEY.md section 8b "Threading").  The handle owns one Engine per device (barcode shard `i` of
// `n_devices`, the same sharding as the one-process-per-GPU path) and one host thread per device for the
// duration of each call.  The per-step exchange of the sums runs over NVLink peer memory mapped by
// cudaDeviceEnablePeerAccess -- same kernels as the multi-process path, no NCCL bootstrap, no IPC.
#pragma once
#include <memory>
#include <thread>

#include "bb_engine.cuh"

namespace bb {

template <typename real> class MultiEngine : public EngineBase {
  public:
    explicit MultiEngine(const bb_desc &d) {
        n_ = d.n_devices;
        int ndev = 0;
        BB_CUDA(cudaGetDeviceCount(&ndev));
        const int base = d.device >= 0 ? d.device : 0;
        if (base + n_ > ndev)
            throw std::runtime_error("n_devices = " + std::to_string(n_) + " from device " + std::to_string(base) +
                                     ": only " + std::to_string(ndev) + " CUDA devices are visible");
        if (n_ > MAX_WORLD) throw std::runtime_error("at most 16 devices per handle");
        eng_.resize(n_);
        parallel([&](int i) {
            bb_desc di = d;
            di.device = base + i; di.rank = i; di.world = n_; di.n_devices = 1;
            eng_[i].reset(new Engine<real>(di));
            eng_[i]->set_raw_elbo(true);
        });
        L = eng_[0]->L;
        L.rank = 0; L.world = 1;
        // peer access between every pair, then every device's exchange buffer to every engine
        parallel([&](int i) {
            for (int j = 0; j < n_; ++j) {
                if (j == i) continue;
                int can = 0;
                BB_CUDA(cudaDeviceCanAccessPeer(&can, eng_[i]->device(), eng_[j]->device()));
                if (!can) throw std::runtime_error("devices " + std::to_string(eng_[i]->device()) + " and " +
                                                   std::to_string(eng_[j]->device()) + " cannot access each other's memory");
                cudaError_t e = cudaDeviceEnablePeerAccess(eng_[j]->device(), 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) BB_CUDA(e);
                cudaGetLastError();
            }
            eng_[i]->alloc_exchange();
        });
        std::vector<void *> bufs(n_);
        for (int i = 0; i < n_; ++i) bufs[i] = eng_[i]->xchg_local();
        parallel([&](int i) { eng_[i]->attach_peers(bufs); });
        refresh();
    }

    void init_params(uint64_t seed) override { parallel([&](int i) { eng_[i]->init_params(seed); }); }
    void set_params(const double *mu, const double *omega) override {
        parallel([&](int i) { eng_[i]->set_params(mu, omega); });
    }
    void get_params(double *mu, double *omega, bool posterior) override {
        gather2(L.D, mu, omega, [&](int i, double *a, double *b) { eng_[i]->get_params(a, b, posterior); });
    }
    void logjoint_grad(const double *x, int K, int eps_is_noise, double *logp, double *grad) override {
        std::vector<std::vector<double>> lp(n_, std::vector<double>(K)), g(n_);
        parallel([&](int i) {
            g[i].resize((size_t)K * L.D);
            eng_[i]->logjoint_grad(x, K, eps_is_noise, lp[i].data(), g[i].data());
        });
        for (int k = 0; k < K; ++k) {
            double s = L.logp_const;
            for (int i = 0; i < n_; ++i) s += lp[i][k];
            logp[k] = s;
        }
        sum_into(grad, g, (size_t)K * L.D);
    }
    void elbo_grad(const double *eps, long long step, double *elbo, double *grad) override {
        std::vector<double> e(n_, 0.0);
        std::vector<std::vector<double>> g(n_);
        parallel([&](int i) {
            if (grad) g[i].resize((size_t)2 * L.D);
            eng_[i]->elbo_grad(eps, step, &e[i], grad ? g[i].data() : nullptr);
        });
        if (elbo) {
            double s = L.logp_const + 0.5 * (double)L.D * (1.0 + std::log(2.0 * M_PI));
            for (int i = 0; i < n_; ++i) s += e[i];
            *elbo = s;
        }
        if (grad) sum_into(grad, g, (size_t)2 * L.D);
    }
    void get_noise(long long step, double *eps) override {
        std::vector<std::vector<double>> g(n_);
        const size_t n = (size_t)L.K * L.D;
        parallel([&](int i) { g[i].resize(n); eng_[i]->get_noise(step, g[i].data()); });
        sum_into(eps, g, n);
    }
    void set_optimizer(const bb_opt &o) override {
        parallel([&](int i) { eng_[i]->set_optimizer(o); });
        refresh();
    }
    void step(int n, double *trace) override {
        std::vector<std::vector<double>> tr(n_);
        parallel([&](int i) {
            if (trace) tr[i].resize(n);
            eng_[i]->step(n, trace ? tr[i].data() : nullptr);
        });
        if (trace) {
            const double cst = L.logp_const + 0.5 * (double)L.D * (1.0 + std::log(2.0 * M_PI));
            for (int s = 0; s < n; ++s) {
                double v = cst;
                for (int i = 0; i < n_; ++i) v += tr[i][s];
                trace[s] = v;
            }
        }
        refresh();
    }
    void step_with_noise(const double *eps) override {
        parallel([&](int i) { eng_[i]->step_with_noise(eps); });
        refresh();
    }
    long long state_size() const override { return eng_[0]->state_size(); }
    void get_state(double *s) override {
        const long long n = state_size();
        std::vector<std::vector<double>> g(n_);
        parallel([&](int i) { g[i].resize(n); eng_[i]->get_state(g[i].data()); });
        const double h0 = g[0][0], h1 = g[0][1];
        sum_into(s, g, (size_t)n);
        s[0] = h0; s[1] = h1;               // step counter and ring slot are replicated, not additive
        // the shared latents and their optimiser state are replicated too: engine 0 reports them, the others 0
    }
    void set_state(const double *s) override {
        parallel([&](int i) { eng_[i]->set_state(s); });
        refresh();
    }
    void set_stream(void *) override {
        throw std::runtime_error("bb_set_stream: a multi-GPU handle runs every device on its own stream");
    }
    void sync() override { parallel([&](int i) { eng_[i]->sync(); }); }
    void time_steps(int n, float *ms_total, float *ms_pass1, float *ms_pass2) override {
        std::vector<float> a(n_), b(n_), c(n_);
        parallel([&](int i) { eng_[i]->time_steps(n, &a[i], &b[i], &c[i]); });
        *ms_total = *std::max_element(a.begin(), a.end());
        if (ms_pass1) *ms_pass1 = *std::max_element(b.begin(), b.end());
        if (ms_pass2) *ms_pass2 = *std::max_element(c.begin(), c.end());
        refresh();
    }
    void derived_fitness(int n, uint64_t seed, double *median, double *sd) override {
        gather2(L.bc_block, median, sd, [&](int i, double *a, double *b) { eng_[i]->derived_fitness(n, seed, a, b); });
    }
    void peer_handle(char *) override { throw std::runtime_error("a multi-GPU handle wires its devices itself"); }
    void peer_attach(const char *, int) override { throw std::runtime_error("a multi-GPU handle wires its devices itself"); }
    void comm_init(const char *) override {}          // the devices of one handle are already connected
    void persist_stats(double out[16]) override { run_on(0, [&] { eng_[0]->persist_stats(out); }); }
    void data_plane(int32_t out[8]) override { eng_[0]->data_plane(out); }
    int n_devices() const { return n_; }

  private:
    template <typename F> void run_on(int i, F &&f) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(eng_[i] ? eng_[i]->device() : prev);
        try { f(); } catch (...) { cudaSetDevice(prev); throw; }
        cudaSetDevice(prev);
    }
    // one host thread per device for the duration of the call; the first exception is rethrown
    template <typename F> void parallel(F &&f) {
        std::vector<std::thread> pool;
        std::vector<std::string> errs(n_);
        std::vector<char> failed(n_, 0);
        for (int i = 0; i < n_; ++i)
            pool.emplace_back([&, i] {
                try {
                    if (eng_[i]) cudaSetDevice(eng_[i]->device());
                    f(i);
                } catch (const std::exception &e) { errs[i] = e.what(); failed[i] = 1; }
            });
        for (auto &t : pool) t.join();
        for (int i = 0; i < n_; ++i)
            if (failed[i]) throw std::runtime_error("device " + std::to_string(i) + ": " + errs[i]);
    }
    static void sum_into(double *dst, const std::vector<std::vector<double>> &parts, size_t n) {
        std::copy(parts[0].begin(), parts[0].begin() + n, dst);
        for (size_t p = 1; p < parts.size(); ++p)
            for (size_t j = 0; j < n; ++j) dst[j] += parts[p][j];
    }
    // every latent is owned (reported non-zero) by exactly one shard: the reference-order vectors add up
    template <typename F> void gather2(long long n, double *x, double *y, F &&f) {
        std::vector<std::vector<double>> a(n_), b(n_);
        parallel([&](int i) { a[i].resize(n); b[i].resize(n); f(i, a[i].data(), b[i].data()); });
        sum_into(x, a, (size_t)n);
        sum_into(y, b, (size_t)n);
    }
    void refresh() {
        launches = 0; alg_bytes = 0.0;
        for (auto &e : eng_) { launches += e->launches; alg_bytes += e->alg_bytes; }
        step_count = eng_[0]->step_count;
    }
    int n_ = 1;
    std::vector<std::unique_ptr<Engine<real>>> eng_;
};

}  // namespace bb
