#pragma once
#include <exception>
#include <boost/exception/all.hpp>
namespace boost {
using std::exception_ptr;
using std::current_exception;
using std::rethrow_exception;
template <class E> inline exception_ptr copy_exception(const E& e) { return std::make_exception_ptr(e); }
}
