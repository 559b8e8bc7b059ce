// Function-pointer bundle of one (real, NT, NE, HIER, SUP) instantiation of the column kernels and the lookup
// implemented by the instantiation units (bb_inst_*.cu / bb_registry.cu).
#pragma once
#include "bb_step_kernel.cuh"

namespace bb {

template <typename real> struct KernelSet {
    void (*pass1)(const P1Args<real>);
    void (*pass2)(const P2Args<real>);        // no ELBO partial sums (the production step)
    void (*pass2_elbo)(const P2Args<real>);   // also accumulates the log-density / entropy partials
    void (*pass2_fused)(const P2Args<real>);  // pass 2 + pass 1 of the next step (non-hierarchical), or nullptr
    // fused step kernel on packs of W samples (bb_step_kernel.cuh): compile-time T and E, non-hierarchical; or nullptr
    StepKernelFn<real> step_w1, step_w2;
};

template <typename real, int NT, int NE, bool HIER, bool SUP> KernelSet<real> make_kernel_set() {
    KernelSet<real> ks;
    ks.pass1 = pass1_kernel<real, NT, NE, HIER, SUP>;
    ks.pass2_elbo = pass2_kernel<real, NT, NE, HIER, SUP, true>;
    // the caller-supplied-noise kernels are the parity path: always with the ELBO terms
    if constexpr (SUP) ks.pass2 = ks.pass2_elbo;
    else ks.pass2 = pass2_kernel<real, NT, NE, HIER, SUP, false>;
    ks.pass2_fused = nullptr;
    if constexpr (!SUP && !HIER) ks.pass2_fused = pass2_kernel<real, NT, NE, HIER, SUP, false, true>;
    ks.step_w1 = nullptr; ks.step_w2 = nullptr;
    if constexpr (!SUP && !HIER && NT > 0 && NE > 0) {
        ks.step_w1 = step_kernel<real, NT, NE, 1>;
        ks.step_w2 = step_kernel<real, NT, NE, 2>;
    }
    return ks;
}

// nt / ne of 0 select the runtime-size kernels
template <typename real> bool lookup_kernels(int nt, int ne, bool hier, bool sup, KernelSet<real> *out);

}  // namespace bb
