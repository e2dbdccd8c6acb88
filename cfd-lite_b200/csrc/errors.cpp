// Error reporting of the C ABI: the reference `stop`s with a message; the library
// returns a code and keeps the message for cfdl_last_error().
#include "cfdl_common.h"
namespace cfdl {
std::string& last_error_ref() { static thread_local std::string s; return s; }
}
extern "C" const char* cfdl_last_error(void) { return cfdl::last_error_ref().c_str(); }
extern "C" int cfdl_version(void) { return 100; }
