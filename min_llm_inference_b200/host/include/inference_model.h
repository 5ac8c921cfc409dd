// drop-in name of the reference header include/inference_model.h; everything lives in mli/compat.hpp
#pragma once
#include "mli/compat.hpp"
