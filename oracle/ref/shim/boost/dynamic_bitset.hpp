#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
namespace boost {
template <typename Block = unsigned long, typename Alloc = void> class dynamic_bitset {
public:
    typedef size_t size_type;
    static const size_type npos = (size_type)-1;
    dynamic_bitset() : n_(0) {}
    explicit dynamic_bitset(size_type n, unsigned long = 0) : n_(n), w_((n + 63) / 64, 0) {}
    size_type size() const { return n_; }
    void resize(size_type n, bool v = false) { w_.resize((n + 63) / 64, v ? ~0ull : 0); n_ = n; }
    void clear() { n_ = 0; w_.clear(); }
    bool test(size_type i) const { return (w_[i >> 6] >> (i & 63)) & 1; }
    bool operator[](size_type i) const { return test(i); }
    struct reference {
        dynamic_bitset& b; size_type i;
        operator bool() const { return b.test(i); }
        reference& operator=(bool v) { b.set(i, v); return *this; }
        reference& operator=(const reference& r) { b.set(i, (bool)r); return *this; }
    };
    reference operator[](size_type i) { return reference{*this, i}; }
    dynamic_bitset& set(size_type i, bool v = true) { if (v) w_[i >> 6] |= 1ull << (i & 63); else w_[i >> 6] &= ~(1ull << (i & 63)); return *this; }
    dynamic_bitset& reset(size_type i) { return set(i, false); }
    dynamic_bitset& reset() { for (auto& x : w_) x = 0; return *this; }
    size_type count() const { size_type c = 0; for (auto x : w_) c += (size_type)__builtin_popcountll(x); return c; }
    bool any() const { for (auto x : w_) if (x) return true; return false; }
    bool none() const { return !any(); }
    size_type find_first() const { return find_from(0); }
    size_type find_next(size_type i) const { return i + 1 >= n_ ? npos : find_from(i + 1); }
private:
    size_type find_from(size_type i) const {
        for (; i < n_; ++i) if (test(i)) return i;
        return npos;
    }
    size_type n_;
    std::vector<uint64_t> w_;
};
}
