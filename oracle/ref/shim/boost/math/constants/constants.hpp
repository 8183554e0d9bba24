#pragma once
namespace boost { namespace math { namespace constants { template <typename T> inline T pi() { return (T)3.14159265358979323846264338327950288L; } } } }
