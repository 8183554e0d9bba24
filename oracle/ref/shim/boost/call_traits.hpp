#pragma once
#include <boost/shared_ptr.hpp>
#include <type_traits>
namespace boost {
template <typename T> struct call_traits {
    typedef T value_type;
    typedef T& reference;
    typedef const T& const_reference;
    typedef typename std::conditional<std::is_scalar<T>::value, const T, const T&>::type param_type;
};
}
