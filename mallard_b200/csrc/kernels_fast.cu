// FAST floating-point mode: compiled with -fmad=true — products feeding sums contract to FMA.
#define MLB_KNS fast
#define MLB_STREAM_KERNELS 1
#include "kernels_impl.cuh"
namespace mlb { const KernelTable * kernels_fast() { return &fast::table; } }
