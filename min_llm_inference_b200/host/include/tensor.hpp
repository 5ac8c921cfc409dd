// drop-in name of the reference header include/tensor.hpp; everything lives in mli/compat.hpp
#pragma once
#include "mli/compat.hpp"
