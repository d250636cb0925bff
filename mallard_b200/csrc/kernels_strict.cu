// STRICT floating-point mode: compiled with -fmad=false (see Makefile) — no FMA contraction anywhere.
#define MLB_KNS strict
#include "kernels_impl.cuh"
namespace mlb { const KernelTable * kernels_strict() { return &strict::table; } }
