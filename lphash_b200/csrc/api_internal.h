// Shared by the translation units behind the C ABI (c_api.cu, build_api.cu): error slot, CUDA error
// propagation, the exception fence every entry point runs inside, grow-only device buffers.
#pragma once
#include <cuda_runtime.h>

#include <new>
#include <stdexcept>
#include <string>

#include "../../include/lphash_b200.h"
#include "lph_image.h"

namespace lphb {

std::string& last_error_slot();  // thread-local; read by lphb_last_error (c_api.cu)

inline int fail(int code, std::string const& msg) {
    last_error_slot() = msg;
    return code;
}

struct CudaError {
    cudaError_t e;
    const char* what;
};
#define CK(expr)                                              \
    do {                                                      \
        cudaError_t e__ = (expr);                             \
        if (e__ != cudaSuccess) throw ::lphb::CudaError{e__, #expr}; \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    uint64_t cap = 0;
    void reserve(uint64_t bytes) {
        if (bytes <= cap) return;
        if (p) CK(cudaFree(p));
        p = nullptr;
        cap = 0;
        uint64_t want = bytes + bytes / 8 + 256;
        CK(cudaMalloc(&p, want));
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        CK(cudaGetDevice(&prev));
        if (prev != dev) CK(cudaSetDevice(dev));
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// nothing may cross the C boundary as an exception
template <class F>
int guarded(F&& body) {
    try {
        return body();
    } catch (CudaError const& e) {
        cudaGetLastError();
        return fail(LPHB_E_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e.e) + " in " + e.what);
    } catch (FormatError const& e) {
        return fail(LPHB_E_FORMAT, e.what());
    } catch (std::bad_alloc const&) {
        return fail(LPHB_E_NOMEM, "out of host memory");
    } catch (std::exception const& e) {
        return fail(LPHB_E_ARG, e.what());
    }
}

}  // namespace lphb
