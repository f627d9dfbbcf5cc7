// sdb_api.cu — library-level entry points of the C ABI (include/sdb.h).
#include "../../include/sdb.h"
#include "common.cuh"

namespace sdb {
char* sdb_error_buffer() {
    static thread_local char buf[kErrLen] = {0};
    return buf;
}
}  // namespace sdb

extern "C" int sdb_version(void) { return 100; }   /* 0.1.0 */
extern "C" const char* sdb_last_error(void) { return sdb::sdb_error_buffer(); }
