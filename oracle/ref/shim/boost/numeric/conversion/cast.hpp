#pragma once
namespace boost { template <typename T, typename S> inline T numeric_cast(S s) { return static_cast<T>(s); } }
