// long8_kernel instantiations for UTF-16 haystacks: class from the high byte / one mixed page (lines8 images), and the SWAR
// images with the class from the high byte or compares on 16-bit lanes.
#include "instances.h"

namespace ndl {
Long8Kernel long8_kernel_utf16(int cm) {
  switch (cm) {
    case kCmHi: return long8_kernel<kCmHi>;
    case kCmMixed: return long8_kernel<kCmMixed>;
#define NDL_QH(k, pl) case cm_swar(k, pl, true): return long8_kernel<cm_swar(k, pl, true)>;
    NDL_QH(4, 1) NDL_QH(4, 2) NDL_QH(2, 1) NDL_QH(2, 2)
#undef NDL_QH
#define NDL_QW(k, pl) case cm_swar_wide(k, pl): return long8_kernel<cm_swar_wide(k, pl)>;
    NDL_QW(4, 1) NDL_QW(4, 2) NDL_QW(4, 3) NDL_QW(2, 1) NDL_QW(2, 2) NDL_QW(2, 3)
#undef NDL_QW
    default: return nullptr;
  }
}
}  // namespace ndl
