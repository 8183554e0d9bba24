// Minimal stand-in for <boost/assert.hpp> (oracle/ref shim: std:: equivalents only, see ../../README.md)
#pragma once
#include <cassert>
#define BOOST_ASSERT(x) assert(x)
#define BOOST_VERIFY(x) ((void)(x))
