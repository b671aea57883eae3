// atk_runtime.hpp -- C++ convenience layer over the C ABI (include/atk_cuda.h) used by the reference-named host
// classes in this directory.  It owns nothing numerical: a process-wide device context, RAII device buffers, and
// the translation of status codes back into the exception types the reference throws.
#pragma once

#include <atk_cuda.h>

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace atk_host {

// ATK status -> the reference's exception types (SURVEY 8(b) error conventions).
inline void check(int status) {
    if (status == ATK_OK) return;
    const std::string msg = atk_last_error();
    switch (status) {
        case ATK_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
        case ATK_ERR_OUT_OF_RANGE: throw std::out_of_range(msg);
        case ATK_ERR_NOMEM: throw std::bad_alloc();
        default: throw std::runtime_error(msg);  // ATK_ERR_RUNTIME, and CUDA / NCCL failures (no CPU fallback)
    }
}

// One context per process, created on first use.  Device: $ATK_DEVICE (default 0).
class Runtime {
public:
    static Runtime& instance() {
        static Runtime rt;
        return rt;
    }
    atk_context_t ctx() {
        if (!ctx_) {
            const char* dev = std::getenv("ATK_DEVICE");
            check(atk_context_create(dev ? std::atoi(dev) : 0, &ctx_));
            if (const char* red = std::getenv("ATK_REDUCTION"))
                check(atk_context_set_reduction(ctx_, std::string(red) == "sequential" ? ATK_REDUCE_SEQUENTIAL : ATK_REDUCE_TREE));
        }
        return ctx_;
    }
    // strictly ordered sums (bit-identical to the reference's std::inner_product) or the fast deterministic tree
    void set_reduction(atk_reduction mode) { check(atk_context_set_reduction(ctx(), mode)); }
    ~Runtime() {
        if (ctx_) atk_context_destroy(ctx_);
    }

private:
    Runtime() = default;
    Runtime(const Runtime&) = delete;
    atk_context_t ctx_ = nullptr;
};

inline atk_context_t context() { return Runtime::instance().ctx(); }

// Device array of doubles, freed on scope exit.
class DeviceArray {
public:
    DeviceArray() = default;
    explicit DeviceArray(size_t n) : n_(n) {
        void* p = nullptr;
        check(atk_malloc(context(), (n ? n : 1) * sizeof(double), &p));
        ptr_ = static_cast<double*>(p);
    }
    DeviceArray(const double* host, size_t n) : DeviceArray(n) {
        if (n) check(atk_memcpy_h2d(context(), ptr_, host, n * sizeof(double)));
    }
    DeviceArray(const DeviceArray&) = delete;
    DeviceArray& operator=(const DeviceArray&) = delete;
    DeviceArray(DeviceArray&& o) noexcept : ptr_(o.ptr_), n_(o.n_) { o.ptr_ = nullptr; }
    ~DeviceArray() {
        if (ptr_) atk_free(context(), ptr_);
    }
    double* get() const { return ptr_; }
    size_t size() const { return n_; }
    void download(double* host) const {
        if (n_) check(atk_memcpy_d2h(context(), host, ptr_, n_ * sizeof(double)));
    }

private:
    double* ptr_ = nullptr;
    size_t n_ = 0;
};

// RAII operator handle.
class OperatorHandle {
public:
    OperatorHandle() = default;
    explicit OperatorHandle(atk_operator_t h) : h_(h) {}
    OperatorHandle(const OperatorHandle&) = delete;
    OperatorHandle& operator=(const OperatorHandle&) = delete;
    OperatorHandle(OperatorHandle&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    OperatorHandle& operator=(OperatorHandle&& o) noexcept {
        reset();
        h_ = o.h_;
        o.h_ = nullptr;
        return *this;
    }
    ~OperatorHandle() { reset(); }
    void reset() {
        if (h_) atk_operator_destroy(h_);
        h_ = nullptr;
    }
    atk_operator_t get() const { return h_; }
    explicit operator bool() const { return h_ != nullptr; }

private:
    atk_operator_t h_ = nullptr;
};

}  // namespace atk_host
