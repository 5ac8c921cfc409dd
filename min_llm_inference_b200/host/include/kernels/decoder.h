// drop-in name of the reference header include/kernels/decoder.h; everything lives in mli/compat.hpp
#pragma once
#include "mli/compat.hpp"
