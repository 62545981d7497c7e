// linesq_kernel instantiations: UTF-16 haystacks (class from the high byte; compares on 16-bit lanes).
#include "instances.h"

namespace ndl {
LinesqKernel linesq_kernel_utf16(int cm) {
  switch (cm) {
#define NDL_QH(k, pl) case cm_swar(k, pl, true): return linesq_kernel<cm_swar(k, pl, true)>;
    NDL_QH(4, 1) NDL_QH(4, 2) NDL_QH(2, 1) NDL_QH(2, 2)
#undef NDL_QH
#define NDL_QW(k, pl) case cm_swar_wide(k, pl): return linesq_kernel<cm_swar_wide(k, pl)>;
    NDL_QW(4, 1) NDL_QW(4, 2) NDL_QW(4, 3) NDL_QW(2, 1) NDL_QW(2, 2) NDL_QW(2, 3)
#undef NDL_QW
    default: return nullptr;
  }
}
}  // namespace ndl
