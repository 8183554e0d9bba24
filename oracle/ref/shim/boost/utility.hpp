#pragma once
#include <boost/noncopyable.hpp>
#include <memory>
namespace boost { using std::addressof; }
