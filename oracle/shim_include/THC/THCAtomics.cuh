// TEST INFRASTRUCTURE (oracle/_ref build only): torch >= 2 no longer ships <THC/THCAtomics.cuh>, which the reference's
// ms_deform_im2col_cuda.cuh includes for atomicAdd in its BACKWARD kernels.  The forward path compiled here needs
// nothing from it; ATen's replacement header keeps the include resolvable without touching the reference source.
#pragma once
#include <ATen/cuda/Atomic.cuh>
