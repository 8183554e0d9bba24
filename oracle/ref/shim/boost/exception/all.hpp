// Minimal stand-in for <boost/exception/all.hpp>: error_info attachment by tag type.
#pragma once
#include <boost/shared_ptr.hpp>
#include <exception>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <typeindex>
#include <typeinfo>
namespace boost {
class exception {
public:
    virtual ~exception() noexcept {}
    mutable std::map<std::type_index, std::shared_ptr<void> > shim_info_;
protected:
    exception() {}
};
template <class Tag, class T> class error_info {
public:
    typedef T value_type;
    error_info(const T& v) : v_(v) {}
    const T& value() const { return v_; }
    T& value() { return v_; }
private:
    T v_;
};
template <class E, class Tag, class T>
inline const E& operator<<(const E& e, const error_info<Tag, T>& info) {
    e.shim_info_[std::type_index(typeid(error_info<Tag, T>))] = std::make_shared<T>(info.value());
    return e;
}
template <class Info, class E>
inline const typename Info::value_type* get_error_info(const E& e) {
    const exception* be = dynamic_cast<const exception*>(&e);
    if (!be) return nullptr;
    auto it = be->shim_info_.find(std::type_index(typeid(Info)));
    return it == be->shim_info_.end() ? nullptr : static_cast<const typename Info::value_type*>(it->second.get());
}
typedef error_info<struct errinfo_file_name_, std::string> errinfo_file_name;
typedef error_info<struct errinfo_errno_, int> errinfo_errno;
typedef error_info<struct errinfo_api_function_, const char*> errinfo_api_function;
template <class E> inline std::string diagnostic_information(const E& e) { return std::string("exception: ") + typeid(e).name(); }
template <class E> [[noreturn]] inline void throw_exception(const E& e) { throw e; }
}  // namespace boost
#define BOOST_THROW_EXCEPTION(x) ::boost::throw_exception(x)
