#pragma once
#include <cstddef>

namespace bn {
// memcpy whose destination is written with non-temporal stores (pinned staging read next by the DMA engine)
void stream_copy(void* dst, const void* src, size_t bytes);
}  // namespace bn
