#pragma once
#include "numeric/conversion/cast.hpp"
namespace boost { template <class T, class S> T polymorphic_downcast(S* s) { return static_cast<T>(s); } }
