// One instantiation unit of the column kernels: compiled once per (BB_REAL, BB_NT)
// by the Makefile (BB_NT = 0 -> runtime number of time points).
#include "bb_kernels.cuh"
#include "bb_registry.h"

#ifndef BB_REAL
#error "BB_REAL must be float or double"
#endif
#ifndef BB_NT
#error "BB_NT must be defined"
#endif

namespace bb {
namespace {
struct Registrar {
    Registrar() {
        register_kernels<BB_REAL>(BB_NT, 1, false, false, make_kernel_set<BB_REAL, BB_NT, 1, false, false>());
        register_kernels<BB_REAL>(BB_NT, 1, false, true, make_kernel_set<BB_REAL, BB_NT, 1, false, true>());
        register_kernels<BB_REAL>(BB_NT, 1, true, false, make_kernel_set<BB_REAL, BB_NT, 1, true, false>());
        register_kernels<BB_REAL>(BB_NT, 1, true, true, make_kernel_set<BB_REAL, BB_NT, 1, true, true>());
        register_kernels<BB_REAL>(BB_NT, 0, false, false, make_kernel_set<BB_REAL, BB_NT, 0, false, false>());
        register_kernels<BB_REAL>(BB_NT, 0, false, true, make_kernel_set<BB_REAL, BB_NT, 0, false, true>());
        register_kernels<BB_REAL>(BB_NT, 0, true, false, make_kernel_set<BB_REAL, BB_NT, 0, true, false>());
        register_kernels<BB_REAL>(BB_NT, 0, true, true, make_kernel_set<BB_REAL, BB_NT, 0, true, true>());
    }
} registrar_instance;
}  // namespace
}  // namespace bb
