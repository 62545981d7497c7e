// long8_kernel instantiations (byte haystacks: the lines8 images and the byte SWAR images).
#include "instances.h"

namespace ndl {
Long8Kernel long8_kernel_bytes(int cm) {
  switch (cm) {
    case kCmBytes: return long8_kernel<kCmBytes>;
    case kCmBytes1: return long8_kernel<kCmBytes1>;
    case kCmBytesH: return long8_kernel<kCmBytesH>;
#define NDL_Q(k, pl, u16) case cm_swar(k, pl, false, u16): return long8_kernel<cm_swar(k, pl, false, u16)>;
    NDL_Q(4, 1, false) NDL_Q(4, 2, false) NDL_Q(4, 3, false)
    NDL_Q(2, 1, false) NDL_Q(2, 2, false) NDL_Q(2, 3, false)
    NDL_Q(2, 1, true) NDL_Q(2, 2, true) NDL_Q(2, 3, true)
    NDL_Q(4, 1, true) NDL_Q(4, 2, true) NDL_Q(4, 3, true)
#undef NDL_Q
    default: return nullptr;
  }
}
}  // namespace ndl
