// Ahead-of-time instantiations of the tree kernel (ps_kernel.cuh) for the common decay trees.
// Compiled once per (PS_AOT_GROUP, PS_EXACT): the exact units are built with -fmad=false.
#include "ps_kernel.cuh"
#include "ps_registry.h"

using psk::Pt;

#ifndef PS_AOT_GROUP
#error "PS_AOT_GROUP must be defined"
#endif

#define PS_GEN(B, D, ...) \
    {#__VA_ARGS__, PS_VARIANT_GENERATE, B, D, PS_EXACT, (const void*)&psk::genbod_kernel<__VA_ARGS__, B != 0, D != 0, PS_EXACT>},
#define PS_REG(...)                                                                                              \
    PS_GEN(0, 0, __VA_ARGS__) PS_GEN(0, 1, __VA_ARGS__) PS_GEN(1, 0, __VA_ARGS__) PS_GEN(1, 1, __VA_ARGS__)      \
    {#__VA_ARGS__, PS_VARIANT_MASK, 0, 0, PS_EXACT, (const void*)&psk::genbod_mask_kernel<__VA_ARGS__, false, PS_EXACT>},   \
    {#__VA_ARGS__, PS_VARIANT_HIST, 0, 0, PS_EXACT, (const void*)&psk::genbod_hist_kernel<__VA_ARGS__, false, PS_EXACT>},

static const PsAotEntry kTable[] = {
#include "ps_aot_trees.inc"
};

namespace {
struct Registrar {
    Registrar() { ps_aot_register(kTable, (int)(sizeof(kTable) / sizeof(kTable[0]))); }
} registrar;
}  // namespace
