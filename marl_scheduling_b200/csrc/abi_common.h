// abi_common.h -- error plumbing shared by the translation units of libmsched.so
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/msched.h"

extern thread_local std::string msched_g_err;  // defined in msched_abi.cu, read by msched_last_error

namespace {

inline int fail(int code, const std::string &msg)
{
    msched_g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(MSCHED_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));    \
    } while (0)

}  // namespace
