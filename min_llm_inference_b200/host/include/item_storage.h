// drop-in name of the reference header include/item_storage.h; everything lives in mli/compat.hpp
#pragma once
#include "mli/compat.hpp"
