// hn_api.h — internal helpers behind the C ABI (include/headnerf_b200.h): error reporting.
#pragma once
#include <cuda_runtime.h>
#include "../../include/headnerf_b200.h"

namespace hn {
int set_error(int code, const char* msg);     // records a thread-local message, returns `code`
int check_launch(const char* what);           // cudaGetLastError() -> 0 or positive cudaError_t
inline int64_t total_samples(int B, int n_rays, int n_samples) { return (int64_t)B * n_rays * n_samples; }
int check_geometry(int B, int n_rays, int n_samples, const char* who);
}  // namespace hn
