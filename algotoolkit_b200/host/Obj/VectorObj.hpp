// VectorObj<T> -- drop-in for the reference's src/Obj/VectorObj.hpp.
//
// Storage is a host std::vector (element() hands out raw pointers, reference :47, so host memory is the source of
// truth between calls).  For T = double every arithmetic member is executed by libatk_cuda through the C ABI
// (atk_dot / atk_nrm2 / atk_vec_*): operands are uploaded, one kernel runs, the result comes back.  Inside the
// solvers (ConjugateGrad, GMRES, Poisson3DSolver) nothing of this is used - they keep all vectors on the device.
// Other element types keep the container interface only; their arithmetic is not provided (device path is fp64).
#pragma once

#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "../atk_runtime.hpp"

template <typename TObj>
class VectorObj {
    std::vector<TObj> data_;

    template <typename U = TObj>
    static constexpr void require_double() {
        static_assert(std::is_same<U, double>::value,
                      "VectorObj arithmetic runs on the GPU and is provided for double only (see DESIGN.md)");
    }
    using binary_fn = int (*)(atk_context_t, int64_t, const double*, const double*, double*);
    using scalar_fn = int (*)(atk_context_t, int64_t, const double*, double, double*);

    VectorObj apply_binary(const VectorObj& other, binary_fn fn) const {
        atk_host::DeviceArray a(data_.data(), data_.size()), b(other.data_.data(), other.data_.size()), out(data_.size());
        atk_host::check(fn(atk_host::context(), (int64_t)data_.size(), a.get(), b.get(), out.get()));
        VectorObj res((int)data_.size());
        out.download(res.data_.data());
        return res;
    }
    void apply_scalar(double s, scalar_fn fn, std::vector<TObj>& dst) const {
        atk_host::DeviceArray a(data_.data(), data_.size()), out(data_.size());
        atk_host::check(fn(atk_host::context(), (int64_t)data_.size(), a.get(), s, out.get()));
        dst.resize(data_.size());
        out.download(dst.data());
    }

public:
    VectorObj() = default;
    explicit VectorObj(int n, TObj value = 0) : data_((size_t)n, value) {}
    VectorObj(const TObj* src, int n) {
        if (!src) throw std::invalid_argument("Null pointer provided to VectorObj constructor.");
        data_.assign(src, src + n);
    }

    // bounds-checked element access (reference :37-45)
    TObj& operator[](int i) {
        if (i < 0 || (size_t)i >= data_.size()) throw std::out_of_range("Index out of range.");
        return data_[(size_t)i];
    }
    const TObj& operator[](int i) const {
        if (i < 0 || (size_t)i >= data_.size()) throw std::out_of_range("Index out of range.");
        return data_[(size_t)i];
    }
    TObj* element() { return data_.data(); }
    const TObj* element() const { return data_.data(); }
    TObj* get_data() const { return const_cast<TObj*>(data_.data()); }
    size_t size() const { return data_.size(); }

    void zero() { data_.assign(data_.size(), TObj()); }
    void resize(size_t n) { data_.resize(n); }
    void resize(size_t n, TObj value) { data_.assign(n, value); }

    // sqrt(sum x^2) (reference :52-54) -> atk_nrm2
    double L2norm() const {
        require_double();
        atk_host::DeviceArray a(data_.data(), data_.size());
        double out = 0.0;
        atk_host::check(atk_nrm2(atk_host::context(), (int64_t)data_.size(), a.get(), &out));
        return out;
    }
    void normalize() {
        const double nrm = L2norm();
        if (nrm == 0.0) throw std::runtime_error("Cannot normalize a zero vector.");
        apply_scalar(nrm, atk_vec_div, data_);
    }

    // x*s, x/s (reference :64-86); division by zero is a runtime_error raised by the library itself
    VectorObj operator*(double s) const {
        require_double();
        VectorObj res;
        apply_scalar(s, atk_vec_scale, res.data_);
        return res;
    }
    VectorObj& operator*=(double s) {
        require_double();
        apply_scalar(s, atk_vec_scale, data_);
        return *this;
    }
    VectorObj operator/(double s) const {
        require_double();
        VectorObj res;
        apply_scalar(s, atk_vec_div, res.data_);
        return res;
    }
    VectorObj& operator/=(double s) {
        require_double();
        apply_scalar(s, atk_vec_div, data_);
        return *this;
    }

    // x+y, x-y (reference :89-102)
    VectorObj operator+(const VectorObj& o) const {
        require_double();
        if (data_.size() != o.data_.size()) throw std::invalid_argument("Vector dimensions do not match.");
        return apply_binary(o, atk_vec_add);
    }
    VectorObj operator-(const VectorObj& o) const {
        require_double();
        if (data_.size() != o.data_.size()) throw std::invalid_argument("Vector dimensions do not match.");
        return apply_binary(o, atk_vec_sub);
    }

    // dot product (reference :105-108) -> atk_dot
    TObj operator*(const VectorObj& o) const {
        require_double();
        if (data_.size() != o.data_.size()) throw std::invalid_argument("Vector dimensions do not match for dot product.");
        atk_host::DeviceArray a(data_.data(), data_.size()), b(o.data_.data(), o.data_.size());
        double out = 0.0;
        atk_host::check(atk_dot(atk_host::context(), (int64_t)data_.size(), a.get(), b.get(), &out));
        return out;
    }
};
