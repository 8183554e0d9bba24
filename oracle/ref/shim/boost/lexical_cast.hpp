#pragma once
#include <boost/shared_ptr.hpp>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
namespace boost {
struct bad_lexical_cast : public std::bad_cast { const char* what() const noexcept override { return "bad lexical cast"; } };
namespace shim_detail {
template <typename S> std::string to_text(const S& s) { std::ostringstream os; os << s; return os.str(); }
inline std::string to_text(const std::string& s) { return s; }
inline std::string to_text(const char* s) { return s; }
inline std::string to_text(char* s) { return s; }
}
template <typename T, typename S>
typename std::enable_if<!std::is_same<T, std::string>::value && !std::is_same<T, const char*>::value, T>::type lexical_cast(const S& s) {
    std::istringstream is(shim_detail::to_text(s));
    T t;
    is >> std::noskipws >> t;
    if (is.fail() || (is.peek() != std::char_traits<char>::eof())) throw bad_lexical_cast();
    return t;
}
template <typename T, typename S>
typename std::enable_if<std::is_same<T, std::string>::value, T>::type lexical_cast(const S& s) { return shim_detail::to_text(s); }
// lexical_cast<const char*>(x): the reference uses it for enum -> name via operator<< (Logger.hh); keep the text alive per thread
template <typename T, typename S>
typename std::enable_if<std::is_same<T, const char*>::value, T>::type lexical_cast(const S& s) {
    static thread_local std::string keep;
    keep = shim_detail::to_text(s);
    return keep.c_str();
}
}  // namespace boost
