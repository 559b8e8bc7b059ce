#pragma once
namespace bb {
template <typename real> struct KernelSet;
template <typename real> void register_kernels(int nt, int ne, bool hier, bool sup, KernelSet<real> ks);
}  // namespace bb
