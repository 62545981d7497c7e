// Which translation unit instantiates which kernel.  The tile kernels are templates over the char mode; compiling all
// of them in one file takes two minutes, so the instantiations are spread over kernels/inst_*.cu (built in parallel by
// the Makefile) and handed out as plain function pointers.  nullptr: that char mode is not instantiated.
#pragma once
#include "lines8.cuh"
#include "long8.cuh"

namespace ndl {

typedef void (*LinesqKernel)(const Lines8Params);
typedef void (*Long8Kernel)(const Long8Params);

LinesqKernel linesq_kernel_bytes4(int cm);  // inst_linesq_bytes4.cu: byte haystacks, 4 chars per lookup (32- and 16-bit entries)
LinesqKernel linesq_kernel_bytes2(int cm);  // inst_linesq_bytes2.cu: byte haystacks, 2 chars per lookup
LinesqKernel linesq_kernel_utf16(int cm);   // inst_linesq_utf16.cu: UTF-16 haystacks (high byte, 16-bit lanes)
Long8Kernel long8_kernel_bytes(int cm);     // inst_long8.cu: byte haystacks
Long8Kernel long8_kernel_utf16(int cm);     // inst_long8_utf16.cu: UTF-16 haystacks

inline LinesqKernel linesq_kernel_for(int cm) {
  if (LinesqKernel k = linesq_kernel_bytes4(cm)) return k;
  if (LinesqKernel k = linesq_kernel_bytes2(cm)) return k;
  return linesq_kernel_utf16(cm);
}

inline Long8Kernel long8_kernel_for(int cm) {
  if (Long8Kernel k = long8_kernel_bytes(cm)) return k;
  return long8_kernel_utf16(cm);
}

}  // namespace ndl
