// FAST floating-point mode: compiled with -fmad=true — products feeding sums contract to FMA.
#define MLB_KNS fast
#include "kernels_impl.cuh"
namespace mlb { const KernelTable * kernels_fast() { return &fast::table; } }
