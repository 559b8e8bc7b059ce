// extern "C" surface declared in include/barbay_b200.h.  Every entry point catches,
// records the message for bb_last_error() and returns a non-zero code: there is no
// CPU fallback -- without a B200 and this library the path fails loudly.
#include <cstring>
#include <memory>
#include <string>

#include "../../include/barbay_b200.h"
#include "bb_multi.cuh"
#include "bb_naive.cuh"

struct bb_handle {
    bb::EngineBase *eng = nullptr;
    std::string err;
};

namespace {
thread_local std::string g_create_error;

// the caller's current device is left as it was found: a handle runs on ITS device whatever is current
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (dev < 0) return;
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename F> int guarded(bb_handle *h, F &&f) {
    if (!h || !h->eng) return 2;
    try {
        DeviceGuard g(h->eng->home_device);
        f(*h->eng);
        h->err.clear();
        return 0;
    } catch (const std::exception &e) {
        h->err = e.what();
        cudaGetLastError();   // clear the sticky-free error state
        return 1;
    }
}
}  // namespace

extern "C" {

int32_t bb_abi_version(void) { return BB_ABI_VERSION; }

int bb_create(const bb_desc *desc, bb_handle **out) {
    if (!desc || !out) { g_create_error = "bb_create: NULL argument"; return 2; }
    *out = nullptr;
    try {
        int ndev = 0;
        cudaError_t ce = cudaGetDeviceCount(&ndev);
        if (ce != cudaSuccess || ndev == 0)
            throw std::runtime_error(std::string("no CUDA device available (") + cudaGetErrorString(ce) +
                                     "); barbay_b200 has no CPU fallback");
        int prev_dev = -1;
        cudaGetDevice(&prev_dev);
        struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
        std::unique_ptr<bb_handle> h(new bb_handle);
        if (desc->n_devices > 1) {
            if (desc->rank != 0 || desc->world != 1)
                throw std::runtime_error("n_devices > 1 shards inside the library: rank / world must be 0 / 1");
            h->eng = desc->dtype == BB_F64 ? bb::make_multi_engine_f64(*desc) : bb::make_multi_engine_f32(*desc);
        } else {
            h->eng = desc->dtype == BB_F64 ? bb::make_engine_f64(*desc) : bb::make_engine_f32(*desc);
        }
        *out = h.release();
        g_create_error.clear();
        return 0;
    } catch (const std::exception &e) {
        g_create_error = e.what();
        cudaGetLastError();
        return 1;
    }
}

int bb_layout_probe(const bb_desc *desc, int32_t *owned, int64_t info[8]) {
    if (!desc || !info) { g_create_error = "bb_layout_probe: NULL argument"; return 2; }
    try {
        bb::Layout L;
        bb::build_layout(*desc, L);
        info[0] = L.D; info[1] = L.n0; info[2] = L.n1; info[3] = L.m0; info[4] = L.m1; info[5] = L.H;
        info[6] = L.hy_gid0; info[7] = L.cpad;
        if (owned) {
            for (long long i = 0; i < L.D; ++i) owned[i] = 0;
            for (const std::vector<int> *m : {&L.map_lam, &L.map_bc, &L.map_hy, &L.map_sh})
                for (int ref : *m)
                    if (ref >= 0) owned[ref] += 1;
        }
        g_create_error.clear();
        return 0;
    } catch (const std::exception &e) {
        g_create_error = e.what();
        return 1;
    }
}

void bb_destroy(bb_handle *h) {
    if (!h) return;
    DeviceGuard g(h->eng ? h->eng->home_device : -1);
    delete h->eng;
    delete h;
}

const char *bb_last_error(const bb_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t bb_n_latent(const bb_handle *h) { return h && h->eng ? h->eng->L.D : -1; }

int bb_init_params(bb_handle *h, uint64_t seed) {
    return guarded(h, [&](bb::EngineBase &e) { e.init_params(seed); });
}
int bb_set_params(bb_handle *h, const double *mu, const double *omega) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (!mu || !omega) throw std::runtime_error("bb_set_params: NULL argument");
        e.set_params(mu, omega);
    });
}
int bb_get_params(bb_handle *h, double *mu, double *omega) {
    return guarded(h, [&](bb::EngineBase &e) { e.get_params(mu, omega, false); });
}
int bb_get_posterior(bb_handle *h, double *m, double *sigma) {
    return guarded(h, [&](bb::EngineBase &e) { e.get_params(m, sigma, true); });
}
int bb_logjoint_grad(bb_handle *h, const double *x, int32_t n_samples, int32_t eps_is_noise, double *logp,
                     double *grad) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (!x || !logp || !grad) throw std::runtime_error("bb_logjoint_grad: NULL argument");
        e.logjoint_grad(x, n_samples, eps_is_noise, logp, grad);
    });
}
int bb_elbo_grad(bb_handle *h, const double *eps, int64_t step, double *elbo, double *grad) {
    return guarded(h, [&](bb::EngineBase &e) { e.elbo_grad(eps, step, elbo, grad); });
}
int bb_get_noise(bb_handle *h, int64_t step, double *eps) {
    return guarded(h, [&](bb::EngineBase &e) { e.get_noise(step, eps); });
}
int bb_set_optimizer(bb_handle *h, const bb_opt *opt) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (!opt) throw std::runtime_error("bb_set_optimizer: NULL argument");
        e.set_optimizer(*opt);
    });
}
int bb_step(bb_handle *h, int32_t n_steps, double *elbo_trace) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (n_steps < 0) throw std::runtime_error("bb_step: n_steps < 0");
        e.step(n_steps, elbo_trace);
    });
}
int bb_step_with_noise(bb_handle *h, const double *eps) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (!eps) throw std::runtime_error("bb_step_with_noise: NULL argument");
        e.step_with_noise(eps);
    });
}
int64_t bb_step_count(const bb_handle *h) { return h && h->eng ? h->eng->step_count : -1; }
int bb_step_until(bb_handle *h, int32_t max_iters, int32_t every, int32_t window, double rel_tol, int32_t *n_done,
                  int32_t *converged, double *elbo_out, int32_t *n_elbo) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (max_iters < 0 || every < 1 || window < 1 || !(rel_tol >= 0.0) || !n_done || !converged)
            throw std::runtime_error("bb_step_until: bad arguments");
        std::vector<double> est;
        int done = 0, conv = 0;
        while (done < max_iters && !conv) {
            const int blk = std::min<int>(every, max_iters - done);
            if (blk > 1) e.step(blk - 1, nullptr);
            double v = 0.0;
            e.step(1, &v);
            done += blk;
            est.push_back(v);
            const size_t n = est.size();
            if (n >= (size_t)2 * window) {
                double a = 0.0, b = 0.0;
                for (int i = 0; i < window; ++i) { a += est[n - 1 - i]; b += est[n - 1 - window - i]; }
                a /= window; b /= window;
                conv = std::fabs(a - b) <= rel_tol * std::fabs(a) ? 1 : 0;
            }
        }
        *n_done = done; *converged = conv;
        if (n_elbo) *n_elbo = (int32_t)est.size();
        if (elbo_out) std::copy(est.begin(), est.end(), elbo_out);
    });
}
int64_t bb_state_size(const bb_handle *h) { return h && h->eng ? h->eng->state_size() : -1; }
int bb_get_state(bb_handle *h, double *s) {
    return guarded(h, [&](bb::EngineBase &e) { e.get_state(s); });
}
int bb_set_state(bb_handle *h, const double *s) {
    return guarded(h, [&](bb::EngineBase &e) { e.set_state(s); });
}
int bb_set_stream(bb_handle *h, void *stream) {
    return guarded(h, [&](bb::EngineBase &e) { e.set_stream(stream); });
}
int bb_sync(bb_handle *h) {
    return guarded(h, [&](bb::EngineBase &e) { e.sync(); });
}
int64_t bb_launch_count(const bb_handle *h) { return h && h->eng ? h->eng->launches : -1; }
double bb_algorithmic_bytes_per_step(const bb_handle *h) { return h && h->eng ? h->eng->alg_bytes : -1.0; }
int bb_time_steps(bb_handle *h, int32_t n_steps, float *ms_total, float *ms_pass1, float *ms_pass2) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (n_steps < 1 || !ms_total) throw std::runtime_error("bb_time_steps: bad arguments");
        e.time_steps(n_steps, ms_total, ms_pass1, ms_pass2);
    });
}
int bb_persist_stats(bb_handle *h, double out[16]) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (!out) throw std::runtime_error("bb_persist_stats: NULL argument");
        e.persist_stats(out);
    });
}
int bb_derived_fitness(bb_handle *h, int32_t n_samples, uint64_t seed, double *median, double *sd) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (!median || !sd) throw std::runtime_error("bb_derived_fitness: NULL argument");
        e.derived_fitness(n_samples, seed, median, sd);
    });
}
int64_t bb_n_derived(const bb_handle *h) { return h && h->eng && h->eng->L.hier ? h->eng->L.bc_block : 0; }
int bb_data_plane(bb_handle *h, int32_t out[8]) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (!out) throw std::runtime_error("bb_data_plane: NULL argument");
        e.data_plane(out);
    });
}
int bb_peer_handle(bb_handle *h, char out[64]) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (!out) throw std::runtime_error("bb_peer_handle: NULL argument");
        e.peer_handle(out);
    });
}
int bb_peer_attach(bb_handle *h, const char *handles, int32_t n) {
    return guarded(h, [&](bb::EngineBase &e) {
        if (!handles) throw std::runtime_error("bb_peer_attach: NULL argument");
        e.peer_attach(handles, n);
    });
}
int bb_naive_prior(const int64_t *bc_count, int32_t n_rep, const int32_t *n_time, int32_t n_neutral, int32_t n_bc,
                   int32_t device, double *s_pop_prior, double *logsig_pop_prior, double *loglam_prior) {
    try {
        int ndev = 0;
        cudaError_t ce = cudaGetDeviceCount(&ndev);
        if (ce != cudaSuccess || ndev == 0)
            throw std::runtime_error(std::string("no CUDA device available (") + cudaGetErrorString(ce) +
                                     "); barbay_b200 has no CPU fallback");
        if (device >= ndev) throw std::runtime_error("bb_naive_prior: no such device");
        DeviceGuard g(device);
        bb::naive_prior_device(bc_count, n_rep, n_time, n_neutral, n_bc, s_pop_prior, logsig_pop_prior, loglam_prior,
                               nullptr);
        g_create_error.clear();
        return 0;
    } catch (const std::exception &e) {
        g_create_error = e.what();
        cudaGetLastError();
        return 1;
    }
}
int bb_comm_unique_id(char id[128]) {
    try {
        bb::NcclApi::UniqueId uid;
        int rc = bb::NcclApi::get().GetUniqueId(&uid);
        if (rc != 0) { g_create_error = "ncclGetUniqueId failed"; return 1; }
        std::memcpy(id, uid.internal, 128);
        return 0;
    } catch (const std::exception &e) {
        g_create_error = e.what();
        return 1;
    }
}
int bb_comm_init(bb_handle *h, const char id[128]) {
    return guarded(h, [&](bb::EngineBase &e) { e.comm_init(id); });
}

}  // extern "C"
