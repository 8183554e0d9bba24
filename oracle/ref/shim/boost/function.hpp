#pragma once
#include <functional>
namespace boost { using std::function; using std::bind; using std::ref; using std::cref; }
