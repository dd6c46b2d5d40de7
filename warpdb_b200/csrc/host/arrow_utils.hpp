// arrow_utils.hpp -- reference: include/arrow_utils.hpp.  Host-side result export, outside the hot path.
#pragma once
#include <cstdint>

#include "arrow_c_abi.h"

// Copies `length` floats into a malloc'd buffer (or a POSIX shared-memory segment "/warpdb_result")
// and describes it as a float32 Arrow array named "result" with no validity bitmap.
void export_to_arrow(const float *data, int64_t length, bool use_shared_memory, ArrowArray *out_array, ArrowSchema *out_schema);
