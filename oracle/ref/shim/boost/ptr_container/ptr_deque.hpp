#pragma once
#include <deque>
#include <memory>
namespace boost {
template <typename T> class ptr_deque {
public:
    typedef std::unique_ptr<T> auto_type;
    void push_back(T* p) { d_.emplace_back(p); }
    void push_front(T* p) { d_.emplace_front(p); }
    auto_type pop_front() { auto_type p = std::move(d_.front()); d_.pop_front(); return p; }
    auto_type pop_back() { auto_type p = std::move(d_.back()); d_.pop_back(); return p; }
    T& front() { return *d_.front(); }
    T& back() { return *d_.back(); }
    T& operator[](size_t i) { return *d_[i]; }
    const T& operator[](size_t i) const { return *d_[i]; }
    size_t size() const { return d_.size(); }
    bool empty() const { return d_.empty(); }
    void clear() { d_.clear(); }
private:
    std::deque<std::unique_ptr<T> > d_;
};
}
