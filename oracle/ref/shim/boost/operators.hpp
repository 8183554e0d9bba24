// Minimal stand-in for <boost/operators.hpp>: only the templates the reference's key types use
// (src/RankSelect.hh:57-68, src/TaggedNum.hh:34-35), with base-class chaining.  A second template
// argument that is an integral type selects the two-type form (T op U); otherwise it is the
// chained base.
#pragma once
#include <boost/shared_ptr.hpp>
#include <type_traits>
namespace boost {
namespace shim_detail {
struct empty_base {};
template <class U, class B> struct pick {                        // (U, B) -> operand type / base type
    typedef typename std::conditional<std::is_integral<U>::value, U, void>::type operand;
    typedef typename std::conditional<std::is_integral<U>::value, B, U>::type base;
};
}  // namespace shim_detail

template <class T, class B = shim_detail::empty_base> struct equality_comparable : B {
    friend bool operator!=(const T& a, const T& b) { return !(a == b); }
};
template <class T, class B = shim_detail::empty_base> struct less_than_comparable : B {
    friend bool operator>(const T& a, const T& b) { return b < a; }
    friend bool operator<=(const T& a, const T& b) { return !(b < a); }
    friend bool operator>=(const T& a, const T& b) { return !(a < b); }
};
template <class T, class B = shim_detail::empty_base> struct incrementable : B {
    friend T operator++(T& x, int) { T tmp(x); ++x; return tmp; }
};
template <class T, class B = shim_detail::empty_base> struct decrementable : B {
    friend T operator--(T& x, int) { T tmp(x); --x; return tmp; }
};

#define GSB_SHIM_BINOP(NAME, OP)                                                                         \
    template <class T, class Operand, class B> struct NAME##_two : B {                                   \
        friend T operator OP(T lhs, const Operand& rhs) { lhs OP## = rhs; return lhs; }                  \
    };                                                                                                   \
    template <class T, class B> struct NAME##_one : B {                                                  \
        friend T operator OP(T lhs, const T& rhs) { lhs OP## = rhs; return lhs; }                        \
    };                                                                                                   \
    template <class T, class U = shim_detail::empty_base, class B = shim_detail::empty_base>             \
    struct NAME : std::conditional<std::is_integral<U>::value, NAME##_two<T, U, B>, NAME##_one<T, U> >::type {};
GSB_SHIM_BINOP(addable, +)
GSB_SHIM_BINOP(subtractable, -)
GSB_SHIM_BINOP(andable, &)
GSB_SHIM_BINOP(orable, |)
GSB_SHIM_BINOP(xorable, ^)
GSB_SHIM_BINOP(left_shiftable, <<)
GSB_SHIM_BINOP(right_shiftable, >>)
#undef GSB_SHIM_BINOP
}  // namespace boost
