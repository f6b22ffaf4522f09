// Stand-in for <cuda_runtime.h> when a thread-per-segment kernel source is compiled as host C++ by the emulation
// harness (tests/emu): everything the kernels use comes from cuda_shim.h.
#pragma once
#include "../cuda_shim.h"
