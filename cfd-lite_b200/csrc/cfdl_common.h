// Shared host-side helpers of libcfdl: error reporting across the C ABI (include/cfdl.h).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <string>
#include "cfdl.h"

namespace cfdl {
std::string& last_error_ref();
inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}
}  // namespace cfdl
