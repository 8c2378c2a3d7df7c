// host_util.h — host-side helpers shared by the translation units behind the C ABI (magpy_b200.cu, comm.cu)
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

struct magpy_b200_comm;

namespace mbh {

// records the message returned by magpy_b200_last_error() on this thread and returns `code`
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
// cudaSetDevice with the library's error reporting (MAGPY_B200_ERR_NO_DEVICE without a usable device)
int select_device(int device);
// comm.cu: in-place all-reduce of a device buffer over the communicator on `stream` (op: MAGPY_B200_COMM_SUM / _MAX)
int comm_allreduce_device(magpy_b200_comm* c, double* dptr, size_t n, int op, cudaStream_t stream);
int comm_device(const magpy_b200_comm* c);

}  // namespace mbh
