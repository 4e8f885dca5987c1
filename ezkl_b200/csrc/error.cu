// error.cu — thread-local last-error string behind b200_last_error() (include/ezkl_b200.h); shared by the product library and
// the test-only debug library.
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace b200 {

static thread_local char tl_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_err, sizeof tl_err, fmt, ap);
    va_end(ap);
}
const char* get_error() { return tl_err; }

}  // namespace b200
