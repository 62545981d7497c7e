// Shared between the host and device halves of the C ABI.
#pragma once
#include <string>

namespace ndl {
// Records the thread-local error message and returns `code`.
int fail(int code, const std::string& msg);
}  // namespace ndl
