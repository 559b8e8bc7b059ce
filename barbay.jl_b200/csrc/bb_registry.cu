// Registry of column-kernel instantiations.  Each bb_inst_* object registers its
// (real, NT, NE, HIER, SUP) kernels at load time; the engine looks up the most
// specialised match and falls back to the runtime-size (NT = 0 / NE = 0) kernels.
#include <vector>

#include "bb_kernel_set.cuh"
#include "bb_registry.h"

namespace bb {

template <typename real> struct Registry {
    struct Entry { int nt, ne; bool hier, sup; KernelSet<real> ks; };
    static std::vector<Entry> &entries() {
        static std::vector<Entry> e;
        return e;
    }
};

template <typename real> void register_kernels(int nt, int ne, bool hier, bool sup, KernelSet<real> ks) {
    Registry<real>::entries().push_back({nt, ne, hier, sup, ks});
}

template <typename real> bool lookup_kernels(int nt, int ne, bool hier, bool sup, KernelSet<real> *out) {
    const int try_nt[2] = {nt, 0};
    const int try_ne[2] = {ne, 0};        // exact environment count if compiled, else the runtime-E kernels
    for (int a : try_nt)
        for (int b : try_ne)
            for (const auto &e : Registry<real>::entries())
                if (e.nt == a && e.ne == b && e.hier == hier && e.sup == sup) { *out = e.ks; return true; }
    return false;
}

template void register_kernels<float>(int, int, bool, bool, KernelSet<float>);
template void register_kernels<double>(int, int, bool, bool, KernelSet<double>);
template bool lookup_kernels<float>(int, int, bool, bool, KernelSet<float> *);
template bool lookup_kernels<double>(int, int, bool, bool, KernelSet<double> *);

}  // namespace bb
