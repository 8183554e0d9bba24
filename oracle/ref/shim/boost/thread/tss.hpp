#pragma once
#include <map>
#include <memory>
namespace boost {
// one object per (thread, instance); enough for src/Profile.hh, which is compiled out anyway
template <typename T> class thread_specific_ptr {
public:
    T* get() const { auto it = slot().find(this); return it == slot().end() ? nullptr : it->second.get(); }
    void reset(T* p = nullptr) { slot()[this].reset(p); }
    T* operator->() const { return get(); }
    T& operator*() const { return *get(); }
private:
    static std::map<const void*, std::unique_ptr<T> >& slot() { static thread_local std::map<const void*, std::unique_ptr<T> > m; return m; }
};
}
