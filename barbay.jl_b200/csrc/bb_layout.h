// Host-side layout: reference latent order (VarInfo order of the Turing models,
// SURVEY.md §8a rows M1-M5) <-> device column space (bb_types.cuh).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/barbay_b200.h"

namespace bb {

struct HostSeg {
    int col0, ncol, rep, nt, neutral;
    uint32_t colid0;
    int sh0;
};

struct Layout {
    // problem
    int model = 0, R = 1, N = 0, M = 0, B = 0, E = 1, G = 0, per = 2;
    bool hier = false;
    std::vector<int> nt;           // [R]
    int tmax = 0, nst = 0;
    std::vector<int> sh0;          // [R] offset of s-bar[r][0] in the s-bar block
    bool as_written = false;       // ragged replicate model: neutral pairing of replicates.jl:599-605
    std::vector<std::vector<int>> env_of_rt;   // [R][kMaxNtDyn] 0-based environment of time point t of replicate r
    int K = 1;
    // shard
    int rank = 0, world = 1;
    int n0 = 0, n1 = 0, m0 = 0, m1 = 0;
    std::vector<int> mperm;        // sorted position -> reference mutant index
    bool perm_identity = true;
    // reference offsets
    long long off_sbar = 0, off_lsbar = 0, off_hyper = 0, off_bc[3] = {0, 0, 0}, off_lam = 0, D = 0;
    std::vector<long long> off_lam_r;
    long long bc_block = 0;        // entries of each barcode-level group
    int H_total = 0, H = 0;
    uint32_t hy_gid0 = 0;
    // column space
    std::vector<HostSeg> segs;
    int cpad = 0, nj = 0;
    std::vector<int> map_lam, map_bc, map_hy, map_sh;   // device slot -> reference index (-1 padding / not owned)
    std::vector<int> cnt;                               // [tmax][cpad]
    std::vector<uint32_t> col_id;                       // [cpad] (only used when !perm_identity)
    std::vector<int> hgroup;                            // [cpad]
    std::vector<int> csr_off, csr_mem;
    // priors, device layout, (mean, 1/var) interleaved as doubles
    bool lam_pr_matrix = false, bc_pr_matrix = false;
    std::vector<double> pr_lam;      // [tmax][cpad][2] if matrix
    std::vector<double> pr_bc;       // [nj][cpad][2] if matrix
    double pr_lam_s[2] = {0, 1};
    double pr_bc_s[3][2] = {{0, 1}, {0, 1}, {0, 1}};
    std::vector<double> pr_sh;       // [2 nst][2]
    std::vector<double> pr_hy;       // [H][2]
    double logp_const = 0.0;         // data / prior normalisation constants of log pi (all shards)
    double n_count_rows = 0;

    int n_local_neutral() const { return n1 - n0; }
    int n_local_mutant() const { return m1 - m0; }
};

// throws std::runtime_error with a message on invalid input
void build_layout(const bb_desc &d, Layout &L);

}  // namespace bb
