// Name -> (offset, count, type) lookup for PbrlModelDesc, so hosts fill the struct by field name.
#pragma once
#include <cstddef>
#include <cstring>

#include "../../include/pbrl.h"

namespace pbrl {
inline int model_desc_field(const char* name, int* offset, int* count, int* is_int) {
#define PBRL_I(f, n) if (!std::strcmp(name, #f)) { *offset = (int)offsetof(PbrlModelDesc, f); *count = n; *is_int = 1; return 0; }
#define PBRL_D(f, n) if (!std::strcmp(name, #f)) { *offset = (int)offsetof(PbrlModelDesc, f); *count = n; *is_int = 0; return 0; }
    PBRL_MODEL_FIELDS(PBRL_I, PBRL_D)
#undef PBRL_I
#undef PBRL_D
    return -1;
}
}  // namespace pbrl
