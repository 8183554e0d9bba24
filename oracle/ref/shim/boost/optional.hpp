#pragma once
#include <boost/assert.hpp>
namespace boost {
struct none_t {};
static const none_t none = none_t();
template <typename T> class optional {
public:
    optional() : has_(false) {}
    optional(none_t) : has_(false) {}
    optional(const T& v) : has_(true), v_(v) {}
    optional& operator=(const T& v) { v_ = v; has_ = true; return *this; }
    optional& operator=(none_t) { has_ = false; return *this; }
    explicit operator bool() const { return has_; }
    bool operator!() const { return !has_; }
    bool is_initialized() const { return has_; }
    T& operator*() { return v_; }
    const T& operator*() const { return v_; }
    T* operator->() { return &v_; }
    const T* operator->() const { return &v_; }
    T& get() { return v_; }
    const T& get() const { return v_; }
    void reset() { has_ = false; }
private:
    bool has_;
    T v_;
};
}
