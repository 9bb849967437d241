/* stand-in: the half-precision shims live in cuda_fake_runtime.h */
#pragma once
#include <cuda_runtime.h>
