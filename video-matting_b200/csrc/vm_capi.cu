// Library-level C ABI: version, thread-local error string, launch checking.
#include "vm_common.cuh"
#include <stdarg.h>
#include <stdio.h>

static thread_local char g_err[512] = "";

void vm_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Launch errors are reported without being consumed (cudaPeekAtLastError): a failure that torch or the caller
// caused earlier on this thread stays visible to them instead of being swallowed - or blamed on this library only.
int vm_check_launch(const char *what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        vm_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return VM_ERR_CUDA;
    }
    return VM_OK;
}

extern "C" int vm_version(void) { return 100; }   /* 0.1.0 */

extern "C" const char *vm_last_error_string(void) { return g_err; }
