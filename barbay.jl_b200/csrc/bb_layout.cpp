// Reference order <-> device column space.  See bb_layout.h / bb_types.cuh.
#include "bb_layout.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <stdexcept>
#include <thread>

namespace bb {

namespace {
constexpr int kMaxNtDyn = 32, kMaxNeDyn = 8, kMaxSeg = 32, kMaxK = 1024;
const double kHalfLog2Pi = 0.5 * std::log(2.0 * M_PI);

[[noreturn]] void fail(const std::string &m) { throw std::runtime_error(m); }

struct PriorView {
    const bb_prior *p;
    const char *name;
    double mean(long long i) const { return p->is_matrix ? p->data[i] : p->data[0]; }
    double sd(long long i) const { return p->is_matrix ? p->data[p->n + i] : p->data[1]; }
};

void check_prior(const bb_prior &p, const char *name, long long rows, bool allow_matrix = true) {
    if (!p.data) fail(std::string("prior ") + name + " is NULL");
    if (p.is_matrix) {
        if (!allow_matrix) fail(std::string("prior ") + name + " must be a 2-vector [mean, std]");
        if (p.n != rows)
            fail(std::string("matrix prior ") + name + " must have " + std::to_string(rows) + " rows, got " +
                 std::to_string(p.n));
        for (long long i = 0; i < rows; ++i)
            if (!(p.data[p.n + i] > 0)) fail(std::string("prior ") + name + ": standard deviations must be > 0");
    } else if (!(p.data[1] > 0)) {
        fail(std::string("prior ") + name + ": standard deviation must be > 0");
    }
}

// sum over rows of (-1/2 log 2 pi - log sd)
double prior_norm_const(const bb_prior &p, long long rows) {
    if (!p.is_matrix) return rows * (-kHalfLog2Pi - std::log(p.data[1]));
    double s = 0.0;
    for (long long i = 0; i < rows; ++i) s += -kHalfLog2Pi - std::log(p.data[p.n + i]);
    return s;
}
}  // namespace

void build_layout(const bb_desc &d, Layout &L) {
    if (d.abi_version != BB_ABI_VERSION) fail("bb_desc.abi_version does not match the library (BB_ABI_VERSION)");
    if (d.model < 0 || d.model > BB_MODEL_MULTIENV_REPLICATE) fail("unknown model variant");
    if (d.dtype != BB_F32 && d.dtype != BB_F64) fail("dtype must be BB_F32 or BB_F64");
    if (d.n_rep < 1 || !d.n_time) fail("n_rep must be >= 1 and n_time non-NULL");
    if (2 * d.n_rep > kMaxSeg) fail("at most 16 replicates are supported");
    if (d.n_neutral < 0 || d.n_bc < 0 || d.n_neutral + d.n_bc < 1) fail("need at least one barcode");
    if (!d.bc_count) fail("bc_count is NULL");
    if (d.n_samples < 1 || d.n_samples > kMaxK) fail("samples_per_step must be in [1, 1024]");
    if (d.world < 1 || d.rank < 0 || d.rank >= d.world) fail("invalid (rank, world)");
    if (d.n_devices < 0) fail("n_devices must be >= 0");

    L.model = d.model;
    L.R = d.n_rep;
    L.N = d.n_neutral;
    L.M = d.n_bc;
    L.B = L.N + L.M;
    L.K = d.n_samples;
    L.rank = d.rank;
    L.world = d.world;
    const bool multienv = d.model == BB_MODEL_MULTIENV || d.model == BB_MODEL_MULTIENV_REPLICATE;
    const bool genotype = d.model == BB_MODEL_GENOTYPE;
    L.hier = d.model == BB_MODEL_REPLICATE || genotype || d.model == BB_MODEL_MULTIENV_REPLICATE;
    L.per = L.hier ? 3 : 2;
    if ((d.model == BB_MODEL_FITNESS_NORMAL || d.model == BB_MODEL_MULTIENV || genotype) && L.R != 1)
        fail("this model takes a single replicate (Matrix{Int64} counts)");
    L.E = multienv ? d.n_env : 1;
    if (L.E < 1 || L.E > kMaxNeDyn) fail("number of environments must be in [1, 8]");
    L.G = genotype ? d.n_geno : 0;

    L.nt.assign(d.n_time, d.n_time + L.R);
    L.tmax = 0;
    L.nst = 0;
    L.sh0.resize(L.R);
    for (int r = 0; r < L.R; ++r) {
        if (L.nt[r] < 2) fail("every replicate needs at least 2 time points");
        if (L.nt[r] > kMaxNtDyn) fail("more than 32 time points per replicate are not supported");
        L.tmax = std::max(L.tmax, L.nt[r]);
        L.sh0[r] = L.nst;
        L.nst += L.nt[r] - 1;
    }
    const bool ragged = *std::min_element(L.nt.begin(), L.nt.end()) != L.tmax;
    if (ragged && multienv && !d.env_per_rep)
        fail("multienv models with unequal time points per replicate need one environment list per replicate (env_per_rep)");
    if (d.env_per_rep && d.model != BB_MODEL_MULTIENV_REPLICATE)
        fail("env_per_rep applies to the multienv x replicate model only");
    // the Vector{Matrix{Int64}} method of the replicate model pairs the neutral ratios as written in
    // replicates.jl:599-605; the multienv x replicate method (…replicates.jl:650-668) pairs them by time
    L.as_written = ragged && d.ragged_as_written && d.model == BB_MODEL_REPLICATE;
    if (L.as_written && d.world > 1)
        fail("the as-written neutral pairing of replicates.jl:599-605 runs on a single shard; pass corrected=true "
             "(ragged_as_written = 0) to shard a fit with unequal time points per replicate");
    L.env_of_rt.assign(L.R, std::vector<int>(kMaxNtDyn, 0));
    if (multienv) {
        if (!d.env_idx) fail("Models with multiple environments require env_idx");
        int off = 0;
        for (int r = 0; r < L.R; ++r) {
            const int32_t *idx = d.env_idx + (d.env_per_rep ? off : 0);
            for (int t = 0; t < L.nt[r]; ++t) {
                if (idx[t] < 1 || idx[t] > L.E) fail("env_idx out of range");
                L.env_of_rt[r][t] = idx[t] - 1;
            }
            off += L.nt[r];
        }
    }
    std::vector<int> g0;   // 0-based genotype index per reference mutant
    if (genotype) {
        if (!d.geno_idx) fail("genotype model requires geno_idx");
        if (L.G < 1) fail("n_geno must be >= 1");
        g0.resize(L.M);
        std::vector<char> seen(L.G, 0);
        for (int m = 0; m < L.M; ++m) {
            if (d.geno_idx[m] < 1 || d.geno_idx[m] > L.G) fail("geno_idx out of range");
            g0[m] = d.geno_idx[m] - 1;
            seen[g0[m]] = 1;
        }
        for (int g = 0; g < L.G; ++g)
            if (!seen[g]) fail("every genotype index in [1, n_geno] must have at least one barcode");
    }

    // ---- reference offsets (VarInfo order)
    L.bc_block = (long long)L.E * L.M * L.R;
    L.H_total = L.hier ? (genotype ? L.G : L.E * L.M) : 0;
    long long off = 0;
    L.off_sbar = off; off += L.nst;
    L.off_lsbar = off; off += L.nst;
    if (L.hier) { L.off_hyper = off; off += L.H_total; }
    for (int kind = 0; kind < L.per; ++kind) { L.off_bc[kind] = off; off += L.bc_block; }
    L.off_lam = off;
    L.off_lam_r.resize(L.R);
    long long ncount = 0;
    for (int r = 0; r < L.R; ++r) {
        L.off_lam_r[r] = off;
        off += (long long)L.nt[r] * L.B;
        ncount += (long long)L.nt[r] * L.B;
    }
    L.D = off;
    L.n_count_rows = (double)ncount;
    if (L.D >= (1LL << 31) - 1) fail("more than 2^31 latent variables are not supported");

    // ---- priors
    check_prior(d.s_pop_prior, "s_pop_prior", L.nst);
    check_prior(d.logsig_pop_prior, "logσ_pop_prior", L.nst);
    check_prior(d.s_bc_prior, "s_bc_prior", L.hier ? L.H_total : L.bc_block);
    check_prior(d.logsig_bc_prior, "logσ_bc_prior", L.bc_block);
    check_prior(d.loglam_prior, "logλ_prior", ncount);
    if (L.hier) check_prior(d.logtau_prior, "logτ_prior", 0, false);

    // ---- shard of the barcode axis
    L.n0 = (int)((long long)L.N * L.rank / L.world);
    L.n1 = (int)((long long)L.N * (L.rank + 1) / L.world);
    L.mperm.resize(L.M);
    std::iota(L.mperm.begin(), L.mperm.end(), 0);
    L.perm_identity = true;
    // genotype model: make every genotype group contiguous -- shards are cut on group boundaries, so theta_g needs no
    // exchange, and on one GPU too the lanes of a warp then share one or two groups: the per-sample fetch of the
    // hyper latent's draw is a broadcast instead of 32 scattered sectors, and the member gather walks consecutive
    // columns (BB_GENO_SORT=0 keeps the caller's order on one GPU: A/B measurements)
    const char *gs = getenv("BB_GENO_SORT");
    if (genotype && (L.world > 1 || !(gs && gs[0] == '0'))) {
        std::stable_sort(L.mperm.begin(), L.mperm.end(), [&](int a, int b) { return g0[a] < g0[b]; });
        for (int p = 0; p < L.M; ++p)
            if (L.mperm[p] != p) { L.perm_identity = false; break; }
        auto boundary = [&](int i) {
            if (i <= 0) return 0;
            if (i >= L.world) return L.M;
            int p = (int)((long long)L.M * i / L.world);
            while (p > 0 && p < L.M && g0[L.mperm[p]] == g0[L.mperm[p - 1]]) ++p;
            return p;
        };
        L.m0 = boundary(L.rank);
        L.m1 = boundary(L.rank + 1);
    } else {
        L.m0 = (int)((long long)L.M * L.rank / L.world);
        L.m1 = (int)((long long)L.M * (L.rank + 1) / L.world);
    }
    const int nN = L.n1 - L.n0, nM = L.m1 - L.m0;

    // ---- hyper latents owned by this shard
    int g_lo = 0;
    if (L.hier) {
        if (genotype) {
            if (nM > 0) {
                g_lo = g0[L.mperm[L.m0]];
                const int g_hi = g0[L.mperm[L.m1 - 1]];
                if (L.world == 1) { g_lo = 0; L.H = L.G; }
                else L.H = g_hi - g_lo + 1;
            } else L.H = 0;
            L.hy_gid0 = (uint32_t)g_lo;
        } else {
            L.H = nM * L.E;
            L.hy_gid0 = (uint32_t)((long long)L.m0 * L.E);
        }
        L.map_hy.resize(L.H);
        for (int h = 0; h < L.H; ++h) L.map_hy[h] = (int)(L.off_hyper + L.hy_gid0 + h);
    }

    // ---- segments
    L.nj = L.per * L.E;
    int col = 0;
    for (int pop = 0; pop < 2; ++pop)
        for (int r = 0; r < L.R; ++r) {
            const int n = pop == 0 ? nN : nM;
            if (n == 0) continue;
            HostSeg s;
            s.col0 = col; s.ncol = n; s.rep = r; s.nt = L.nt[r]; s.neutral = pop == 0;
            s.colid0 = (uint32_t)((long long)r * L.B + (pop == 0 ? L.n0 : L.N + L.m0));
            s.sh0 = L.sh0[r];
            L.segs.push_back(s);
            col += (n + 31) / 32 * 32;
        }
    L.cpad = std::max(col, 32);
    // the kernels address latents with 32-bit element offsets (row * cpad + column)
    if ((long long)std::max(L.tmax, L.nj) * L.cpad >= (1LL << 31))
        throw std::runtime_error("problem too large for one device: (time points x columns) must stay below 2^31 per shard");

    // ---- maps, counts, ids
    L.map_lam.assign((size_t)L.tmax * L.cpad, -1);
    L.map_bc.assign((size_t)L.nj * L.cpad, -1);
    L.cnt.assign((size_t)L.tmax * L.cpad, 0);
    if (!L.perm_identity) L.col_id.assign(L.cpad, 0);
    if (L.hier) L.hgroup.assign(L.cpad, 0);
    std::vector<long long> cnt_off(L.R);
    {
        long long o = 0;
        for (int r = 0; r < L.R; ++r) { cnt_off[r] = o; o += (long long)L.nt[r] * L.B; }
    }
    std::vector<std::vector<int>> members(L.hier ? L.H : 0);
    for (const HostSeg &s : L.segs) {
        for (int i = 0; i < s.ncol; ++i) {
            const int c = s.col0 + i;
            const int mref = s.neutral ? -1 : L.mperm[L.m0 + i];
            const int b = s.neutral ? L.n0 + i : L.N + mref;
            if (!L.perm_identity) L.col_id[c] = (uint32_t)((long long)s.rep * L.B + b);
            for (int t = 0; t < s.nt; ++t) {
                L.map_lam[(size_t)t * L.cpad + c] = (int)(L.off_lam_r[s.rep] + (long long)b * s.nt + t);
                const int64_t v = d.bc_count[cnt_off[s.rep] + (long long)b * s.nt + t];
                if (v < 0) fail("negative barcode count");
                if (v > 2147483647LL) fail("barcode counts above 2^31-1 are not supported");
                L.cnt[(size_t)t * L.cpad + c] = (int)v;
            }
            if (s.neutral) continue;
            for (int e = 0; e < L.E; ++e)
                for (int kind = 0; kind < L.per; ++kind)
                    L.map_bc[(size_t)(L.per * e + kind) * L.cpad + c] =
                        (int)(L.off_bc[kind] + e + (long long)L.E * (mref + (long long)L.M * s.rep));
            if (L.hier) {
                const int hb = genotype ? g0[mref] - g_lo : i * L.E;
                L.hgroup[c] = hb;
                for (int e = 0; e < L.E; ++e) members[hb + e].push_back(e * L.cpad + c);
            }
        }
    }
    if (L.hier) {
        L.csr_off.assign(L.H + 1, 0);
        for (int h = 0; h < L.H; ++h) L.csr_off[h + 1] = L.csr_off[h] + (int)members[h].size();
        L.csr_mem.reserve(L.csr_off[L.H]);
        for (int h = 0; h < L.H; ++h) L.csr_mem.insert(L.csr_mem.end(), members[h].begin(), members[h].end());
    }
    L.map_sh.resize(2 * L.nst);
    for (int i = 0; i < 2 * L.nst; ++i) L.map_sh[i] = L.rank == 0 ? i : -1;   // replicated; rank 0 reports them

    // ---- priors in device layout
    auto pack = [](double mean, double sd, double *out) { out[0] = mean; out[1] = 1.0 / (sd * sd); };
    L.pr_sh.resize((size_t)4 * L.nst);
    for (int i = 0; i < L.nst; ++i) {
        PriorView a{&d.s_pop_prior, ""}, b{&d.logsig_pop_prior, ""};
        pack(a.mean(i), a.sd(i), &L.pr_sh[2 * (size_t)i]);
        pack(b.mean(i), b.sd(i), &L.pr_sh[2 * (size_t)(L.nst + i)]);
    }
    const bb_prior unit_prior_dummy{nullptr, 0, 0};
    (void)unit_prior_dummy;
    const double unit[2] = {0.0, 1.0};
    const bb_prior unit_prior{unit, 0, 0};
    const bb_prior *kind_prior[3];
    if (L.hier) { kind_prior[0] = &unit_prior; kind_prior[1] = &d.logtau_prior; kind_prior[2] = &d.logsig_bc_prior; }
    else { kind_prior[0] = &d.s_bc_prior; kind_prior[1] = &d.logsig_bc_prior; kind_prior[2] = nullptr; }
    L.bc_pr_matrix = false;
    for (int kind = 0; kind < L.per; ++kind) {
        PriorView v{kind_prior[kind], ""};
        if (kind_prior[kind]->is_matrix) L.bc_pr_matrix = true;
        else pack(v.mean(0), v.sd(0), L.pr_bc_s[kind]);
    }
    if (L.bc_pr_matrix) {
        L.pr_bc.assign((size_t)2 * L.nj * L.cpad, 0.0);
        for (size_t s = 0; s < (size_t)L.nj * L.cpad; ++s) {
            const int j = (int)(s / L.cpad), kind = j % L.per;
            PriorView v{kind_prior[kind], ""};
            const int ref = L.map_bc[s];
            if (ref < 0) { L.pr_bc[2 * s] = 0.0; L.pr_bc[2 * s + 1] = 1.0; continue; }
            const long long row = ref - L.off_bc[kind];
            pack(v.mean(row), v.sd(row), &L.pr_bc[2 * s]);
        }
    }
    L.lam_pr_matrix = d.loglam_prior.is_matrix != 0;
    {
        PriorView v{&d.loglam_prior, ""};
        if (L.lam_pr_matrix) {
            L.pr_lam.assign((size_t)2 * L.tmax * L.cpad, 0.0);
            for (size_t s = 0; s < (size_t)L.tmax * L.cpad; ++s) {
                const int ref = L.map_lam[s];
                if (ref < 0) { L.pr_lam[2 * s] = 0.0; L.pr_lam[2 * s + 1] = 1.0; continue; }
                pack(v.mean(ref - L.off_lam), v.sd(ref - L.off_lam), &L.pr_lam[2 * s]);
            }
        } else pack(v.mean(0), v.sd(0), L.pr_lam_s);
    }
    if (L.hier) {
        PriorView v{&d.s_bc_prior, ""};
        L.pr_hy.resize((size_t)2 * L.H);
        for (int h = 0; h < L.H; ++h) pack(v.mean(L.hy_gid0 + h), v.sd(L.hy_gid0 + h), &L.pr_hy[2 * (size_t)h]);
    }

    // ---- constants of log pi (identical on every shard)
    double cst = 0.0;
    cst += prior_norm_const(d.s_pop_prior, L.nst) + prior_norm_const(d.logsig_pop_prior, L.nst);
    if (L.hier) {
        cst += prior_norm_const(d.s_bc_prior, L.H_total);
        cst += L.bc_block * (-kHalfLog2Pi);                                   // theta-tilde ~ N(0, 1)
        cst += prior_norm_const(d.logtau_prior, L.bc_block);
        cst += prior_norm_const(d.logsig_bc_prior, L.bc_block);
    } else {
        cst += prior_norm_const(d.s_bc_prior, L.bc_block) + prior_norm_const(d.logsig_bc_prior, L.bc_block);
    }
    cst += prior_norm_const(d.loglam_prior, ncount);
    // sum of lgamma(r + 1) over every count: 64 fixed chunks (so the rounding does not depend on the machine),
    // spread over the host threads -- at 5 * 10^6 counts this loop was the largest part of bb_create
    double lg = 0.0;
    {
        constexpr int kChunks = 64;
        double partial[kChunks] = {0.0};
        const int nthr = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        // read counts are small integers: lgamma(r + 1) of r < 4096 comes from a table of the same libm values
        constexpr int kTab = 4096;
        std::vector<double> tab(kTab);
        {
            int sign = 0;
            for (int r = 0; r < kTab; ++r) tab[r] = lgamma_r((double)r + 1.0, &sign);
        }
        auto work = [&](int w) {
            int sign = 0;
            for (int ch = w; ch < kChunks; ch += nthr) {
                const long long i0 = ncount * ch / kChunks, i1 = ncount * (ch + 1) / kChunks;
                double acc = 0.0;
                for (long long i = i0; i < i1; ++i) {
                    const int64_t r = d.bc_count[i];
                    acc += (r >= 0 && r < kTab) ? tab[r] : lgamma_r((double)r + 1.0, &sign);
                }
                partial[ch] = acc;
            }
        };
        if (ncount < 100000 || nthr == 1) {
            for (int w = 0; w < nthr; ++w) work(w);
        } else {
            std::vector<std::thread> pool;
            for (int w = 0; w < nthr; ++w) pool.emplace_back(work, w);
            for (auto &t : pool) t.join();
        }
        for (int ch = 0; ch < kChunks; ++ch) lg += partial[ch];
    }
    cst -= lg;
    long long nratio = 0;
    for (int r = 0; r < L.R; ++r) nratio += (long long)(L.nt[r] - 1) * L.B;
    cst -= kHalfLog2Pi * (double)nratio;
    L.logp_const = cst;
}

}  // namespace bb
