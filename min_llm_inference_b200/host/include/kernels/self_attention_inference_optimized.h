// drop-in name of the reference header include/kernels/self_attention_inference_optimized.h; everything lives in mli/compat.hpp
#pragma once
#include "mli/compat.hpp"
