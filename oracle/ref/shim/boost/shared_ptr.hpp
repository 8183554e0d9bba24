#pragma once
#include <functional>
#include <boost/assert.hpp>
#include <boost/static_assert.hpp>
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
}
