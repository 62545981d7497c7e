// linesq_kernel instantiations: byte haystacks, 4 chars per transition lookup.
#include "instances.h"

namespace ndl {
LinesqKernel linesq_kernel_bytes4(int cm) {
  switch (cm) {
#define NDL_Q(pl, u16) case cm_swar(4, pl, false, u16): return linesq_kernel<cm_swar(4, pl, false, u16)>;
    NDL_Q(1, false) NDL_Q(2, false) NDL_Q(3, false) NDL_Q(1, true) NDL_Q(2, true) NDL_Q(3, true)
#undef NDL_Q
    default: return nullptr;
  }
}
}  // namespace ndl
