// One instantiation unit of the column kernels: compiled once per (BB_REAL, BB_NT)
// by the Makefile (BB_NT = 0 -> runtime number of time points).  BB_NE_SEL (optional) restricts the unit
// to the one-environment (1) or runtime-environment (0) kernels: the runtime-T units are the slowest to
// compile and are split in two so the build parallelises.
#include "bb_kernel_set.cuh"
#include "bb_registry.h"

#ifndef BB_REAL
#error "BB_REAL must be float or double"
#endif
#ifndef BB_NT
#error "BB_NT must be defined"
#endif
#ifndef BB_NE_SEL
#define BB_NE_SEL -1
#endif

namespace bb {
namespace {
struct Registrar {
    Registrar() {
#if BB_NE_SEL > 1
        // multi-environment production kernels with a compile-time number of environments (no caller-supplied
        // noise, no hierarchy): the runtime-E kernels size their per-environment state for 8 environments and spill
        register_kernels<BB_REAL>(BB_NT, BB_NE_SEL, false, false, make_kernel_set<BB_REAL, BB_NT, BB_NE_SEL, false, false>());
#else
#if BB_NE_SEL != 0
        register_kernels<BB_REAL>(BB_NT, 1, false, false, make_kernel_set<BB_REAL, BB_NT, 1, false, false>());
        register_kernels<BB_REAL>(BB_NT, 1, false, true, make_kernel_set<BB_REAL, BB_NT, 1, false, true>());
        register_kernels<BB_REAL>(BB_NT, 1, true, false, make_kernel_set<BB_REAL, BB_NT, 1, true, false>());
        register_kernels<BB_REAL>(BB_NT, 1, true, true, make_kernel_set<BB_REAL, BB_NT, 1, true, true>());
#endif
#if BB_NE_SEL != 1
        register_kernels<BB_REAL>(BB_NT, 0, false, false, make_kernel_set<BB_REAL, BB_NT, 0, false, false>());
        register_kernels<BB_REAL>(BB_NT, 0, false, true, make_kernel_set<BB_REAL, BB_NT, 0, false, true>());
        register_kernels<BB_REAL>(BB_NT, 0, true, false, make_kernel_set<BB_REAL, BB_NT, 0, true, false>());
        register_kernels<BB_REAL>(BB_NT, 0, true, true, make_kernel_set<BB_REAL, BB_NT, 0, true, true>());
#endif
#endif
    }
} registrar_instance;
}  // namespace
}  // namespace bb
