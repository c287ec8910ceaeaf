#include "util.h"
#include <stdexcept>
#include <vector>

namespace evq {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

const char* last_error() { return g_last_error.c_str(); }

void fail(int status, const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error{status, buf};
}

}  // namespace evq
