#pragma once
#include <stdint.h>
#include <vector>
#include "../../include/herro_b200.h"
namespace hb {
// returns 0 ok (possibly no windows), -1 on input the reference would panic on
int host_extract_windows(const hb_overlap& o, uint32_t overlap_idx, uint32_t W, uint32_t n_windows,
                         std::vector<hb_overlap_window>& out);
}
